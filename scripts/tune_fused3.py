"""GPU scratch tool: fused kernel latency/throughput on simple unit populations."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, ViterbiEngine  # noqa: E402
from tests import synth  # noqa: E402

dev = torch.device("cuda:0")
eng = ViterbiEngine(dev)


def time_plan(T, trs, label, reps=10):
    rng = np.random.default_rng(1)
    means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, 48, int(t))
                      for tr, t in zip(trs, T)])
    T = np.asarray(T)
    logp = torch.log_softmax(torch.randn(int(T.sum()), 48, device=dev), dim=1).contiguous()
    plan = AlignPlan(T, [[list(map(int, t))] for t in trs], 48, device=dev, len_params=poisson_params(means),
                     long_K=10**6)
    res = {}
    for mode in ("fused", "split"):
        for _ in range(3):
            eng.run(plan, logp, seg0_f32=True, mode=mode, write_bs=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.run(plan, logp, seg0_f32=True, mode=mode, write_bs=False)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / reps * 1e3
    print(f"{label:40s} units={len(T):5d} fused={res['fused']:8.1f} us  split={res['split']:8.1f} us", flush=True)


rng = np.random.default_rng(0)
mk = lambda n: rng.integers(0, 48, n)
time_plan([10000], [mk(6)], "1 unit T=10000 N=6")
time_plan([10000], [mk(12)], "1 unit T=10000 N=12")
time_plan([2000], [mk(6)], "1 unit T=2000 N=6")
time_plan([10000] * 148, [mk(6) for _ in range(148)], "148 units T=10000 N=6")
time_plan([10000] * 444, [mk(6) for _ in range(444)], "444 units T=10000 N=6")
time_plan([10000] * 444, [mk(12) for _ in range(444)], "444 units T=10000 N=12")
time_plan([2000] * 1776, [mk(6) for _ in range(1776)], "1776 units T=2000 N=6")
time_plan([2000] * 1776, [mk(12) for _ in range(1776)], "1776 units T=2000 N=12")
time_plan([300] * 8880, [mk(4) for _ in range(8880)], "8880 units T=300 N=4")
T, trs, _ = bench.make_split(0)
time_plan(T, trs, "c2 split")
o = np.argsort(-T)
time_plan(T[o][200:], [trs[i] for i in o[200:]], "c2 without the 200 longest")
time_plan(np.minimum(T, 4000), trs, "c2 with T clipped at 4000")
