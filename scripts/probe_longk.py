"""GPU scratch tool: c2 with the longest videos in a separate wide launch (a warp per segment)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
logp = bench.device_logp(T, trs, 0, dev)
eng = ViterbiEngine(dev)
frames = int(T.sum())
ref = None
for long_K in (0, None, 320, 310, 300, 293, 283, 266):
    plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means), long_K=long_K)
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    torch.cuda.synchronize()
    out = eng.fetch(plan)
    if ref is None:
        ref = out
    ok = all(np.array_equal(out[k], ref[k]) for k in ("score", "labels", "seg_blocks"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    e1.record(); torch.cuda.synchronize()
    print(f"long_K={long_K} n_long={plan.n_long}  {e0.elapsed_time(e1)/30*1e3:7.1f} us exact={ok}", flush=True)
