"""GPU scratch tool: times the lane-per-segment DP kernel on c2 (after a completed scan)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, ViterbiEngine  # noqa: E402

dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
logp = bench.device_logp(T, trs, 0, dev)
eng = ViterbiEngine(dev)
plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means))
print("units", plan.U, "lane warps", plan.n_lane_warps, "lane fill",
      float((plan.lane_unit >= 0).mean()), flush=True)
eng.run(plan, logp, seg0_f32=True, mode="fused")
torch.cuda.synchronize()
ref = eng.fetch(plan, want_bp=True)
frames = int(T.sum())
for mode in ("split", "lanes", "fused"):
    mid = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode=mode)
    torch.cuda.synchronize()
    out = eng.fetch(plan, want_bp=True)
    ok = all(np.array_equal(out[k], ref[k]) for k in ("score", "labels", "seg_blocks", "final_j", "bp", "status"))
    tot, tail = [], []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(plan, logp, seg0_f32=True, mode=mode, mid_event=mid)
        e1.record()
        torch.cuda.synchronize()
        tot.append(e0.elapsed_time(e1)); tail.append(mid.elapsed_time(e1))
    print(f"{mode:6s} total {np.median(tot)*1e3:7.1f} us   after-scan part {np.median(tail)*1e3:7.1f} us   exact={ok}", flush=True)
