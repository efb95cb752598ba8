"""GPU scratch tool: the longest videos of c2 aligned alone (uncontended critical path)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
eng = ViterbiEngine(dev)
order = np.argsort(-T)
for nlong in (1, 20, 148, 296, 592):
    idx = order[:nlong]
    Ts = T[idx]; trl = [trs[i] for i in idx]; ms_ = means[idx]
    logp = bench.device_logp(Ts, trl, 0, dev)
    plan = AlignPlan(Ts, [[t.tolist()] for t in trl], 48, device=dev, len_params=poisson_params(ms_))
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"longest {nlong:4d}: maxT={Ts.max()} N(longest)={len(trl[0])} minT={Ts.min()}  {us:7.1f} us  ({us*1e3/(Ts.max()//30):6.1f} ns/step of the longest)", flush=True)
