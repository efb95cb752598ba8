"""GPU scratch tool: fused alignment on c2 as a function of the long-video threshold."""
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
frames = int(T.sum())
ref = None
for long_K in (10**6, 300, 250, 200, 150, 120, 100, 80, 60):
    plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means), long_K=long_K)
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out = eng.fetch(plan)
    if ref is None:
        ref = out
    ok = np.array_equal(out["labels"], ref["labels"]) and np.array_equal(out["score"], ref["score"])
    print(f"long_K={long_K:8d} n_long={plan.n_long:5d}  {ms*1e3:7.1f} us  {frames/ms/1e6:7.2f} Gframes/s  exact={ok}", flush=True)
