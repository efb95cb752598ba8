"""GPU: timings of the other BASELINE.json configs (c1, c3 subset, c4) -- parity for these shapes is in
tests/test_viterbi_gpu.py; this only records numbers for DESIGN.md."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import PoissonModel, SingleTranscriptGrammar  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, Viterbi, ViterbiEngine  # noqa: E402
from tests import synth  # noqa: E402

dev = torch.device("cuda:0")
eng = ViterbiEngine(dev)
res = {}


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# c1: one Breakfast-shaped video through the reference call signature (host numpy in, lists out)
rng = np.random.default_rng(0)
tr = [0, 5, 7, 5, 12, 0]
lp, _ = synth.planted_logp(rng, 2000, 48, tr, np.float32)
means = synth.class_means(rng.dirichlet(5 * np.ones(6)).astype(np.float32), tr, 48, 2000)
dec = Viterbi(SingleTranscriptGrammar(tr, 48), PoissonModel(means), frame_sampling=30, device=dev)
for _ in range(3):
    dec.decode(lp)
t0 = time.perf_counter()
for _ in range(20):
    dec.grammar = SingleTranscriptGrammar(tr, 48)
    dec.length_model = PoissonModel(means)
    dec.decode(lp)
res["c1_drop_in_decode_ms"] = (time.perf_counter() - t0) / 20 * 1e3
plan = AlignPlan([2000], [[tr]], 48, device=dev, len_params=poisson_params(means)[None])
dlp = torch.from_numpy(lp).to(dev)
res["c1_kernel_only_us"] = timeit(lambda: eng.run(plan, dlp, seg0_f32=True, write_bs=False)) * 1e3

# c3 (subset): 256 videos x 64 candidate transcripts, two-kernel path + select + winner labels
T, trs, _ = bench.make_split(0)
nv = 256
Ts = T[:nv]
cands, mlist = [], []
r2 = np.random.default_rng(5)
for v in range(nv):
    K = int(Ts[v]) // 30
    cl = synth.random_edits(r2, trs[v], 48, 64, max(2, -(-K // 66)), min(30, K))
    cands.append(cl)
    mlist.append(synth.class_means(r2.dirichlet(np.ones(len(trs[v]))).astype(np.float32), trs[v], 48, int(Ts[v])))
logp = bench.device_logp(Ts, trs[:nv], 0, dev)
plan3 = AlignPlan(Ts, cands, 48, device=dev, len_params=poisson_params(np.stack(mlist)), labels="best")
ms = timeit(lambda: eng.run(plan3, logp, seg0_f32=True), n=5)
res["c3_subset"] = {"videos": nv, "candidates": 64, "units": plan3.U, "max_N": plan3.max_N, "ms": ms,
                    "mode": eng.last_mode,
                    "aligned_frames_per_s": plan3.aligned_frames / (ms * 1e-3), "ctas": plan3.n_cta, "wpc": plan3.wpc}
eng.run(plan3, logp, seg0_f32=True, mode="split")
torch.cuda.synchronize()
ref3 = eng.fetch(plan3)
for m3 in ("split", "lanes"):
    if m3 == "lanes" and plan3.lane_unit is None:
        continue
    ms = timeit(lambda: eng.run(plan3, logp, seg0_f32=True, mode=m3), n=5)
    o3 = eng.fetch(plan3)
    res["c3_subset_" + m3] = {"ms": ms, "aligned_frames_per_s": plan3.aligned_frames / (ms * 1e-3),
                              "lane_warps": plan3.n_lane_warps,
                              "same_as_split": bool(all(np.array_equal(o3[k], ref3[k]) for k in ("score", "labels", "best", "seg_blocks")))}

# c4: one long video, T = 40000, C = 100, N = 60
r4 = np.random.default_rng(4)
tr4 = r4.permutation(100)[:60].tolist()
lp4, _ = synth.planted_logp(r4, 40000, 100, tr4, np.float32)
m4 = synth.class_means(r4.dirichlet(5 * np.ones(60)).astype(np.float32), tr4, 100, 40000)
plan4 = AlignPlan([40000], [[tr4]], 100, device=dev, len_params=poisson_params(m4)[None])
d4 = torch.from_numpy(lp4).to(dev)
res["c4_long_video_us"] = timeit(lambda: eng.run(plan4, d4, seg0_f32=True, write_bs=False), n=5) * 1e3
res["c4_mode"] = eng.last_mode
print(json.dumps(res))
