"""GPU scratch tool: is the fused alignment bound by the longest video's serial DP chain or by
throughput?  Times c2 as is, and c2-like batches with the same total frames whose video lengths are
capped (long videos cut into pieces), so the critical path shrinks while the work stays."""
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
eng = ViterbiEngine(dev)
frames = int(T.sum())
logp = bench.device_logp(T, trs, 0, dev)


def timeit(plan, n=30):
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for cap in (0, 6000, 4000, 3000, 2000, 1000):
    Tc, trc, mc = [], [], []
    for t, tr, m in zip(T, trs, means):
        t = int(t)
        if cap and t > cap:
            n = -(-t // cap)
            base = t // n
            parts = [base + (1 if i < t - base * n else 0) for i in range(n)]
        else:
            parts = [t]
        for p in parts:
            Tc.append(p); trc.append(tr); mc.append(m)
    Tc = np.asarray(Tc, dtype=np.int64)
    assert Tc.sum() == frames
    plan = AlignPlan(Tc, [[t.tolist()] for t in trc], 48, device=dev, len_params=poisson_params(np.stack(mc)))
    ms = timeit(plan)
    print(f"cap={cap:5d} videos={len(Tc):5d} maxT={Tc.max():5d}  {ms*1e3:7.1f} us  {frames/ms/1e6:7.2f} Gframes/s", flush=True)
