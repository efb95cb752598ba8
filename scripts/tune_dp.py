"""GPU scratch tool: times the DP kernel alone on a few synthetic unit populations."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import _lib  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, ViterbiEngine  # noqa: E402
from tests import synth  # noqa: E402

dev = torch.device("cuda:0")
eng = ViterbiEngine(dev)
lib = _lib.lib()


def time_plan(T, trs, label, reps=20):
    rng = np.random.default_rng(1)
    means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, 48, int(t))
                      for tr, t in zip(trs, T)])
    T = np.asarray(T)
    logp = torch.log_softmax(torch.randn(int(T.sum()), 48, device=dev), dim=1).contiguous()
    plan = AlignPlan(T, [[list(map(int, t))] for t in trs], 48, device=dev, len_params=poisson_params(means))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True)
    torch.cuda.synchronize()
    sc, dp = [], []
    for _ in range(reps):
        ev[0].record()
        eng.run(plan, logp, seg0_f32=True, mid_event=ev[1])
        ev[2].record()
        torch.cuda.synchronize()
        sc.append(ev[0].elapsed_time(ev[1])); dp.append(ev[1].elapsed_time(ev[2]))
    K = T // 30
    steps = int(K.max())
    print(f"{label:44s} units={len(T):5d} ctas={plan.n_cta:5d} wpc={plan.wpc} maxK={steps:4d} "
          f"scan={np.median(sc)*1e3:7.1f}us dp={np.median(dp)*1e3:7.1f}us  dp/maxK={np.median(dp)*1e6/steps:7.1f} ns/step", flush=True)


rng = np.random.default_rng(0)
mk = lambda n: rng.integers(0, 48, n)
if len(sys.argv) > 1 and sys.argv[1] == "persm":
    time_plan([10000] * 148, [mk(6) for _ in range(148)], "148 units T=10000 N=6", reps=2)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "perwarp":
    time_plan([10000] * 592, [mk(6) for _ in range(592)], "592 units T=10000 N=6", reps=2)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "single":
    time_plan([10000], [mk(6)], "1 unit T=10000 N=6 (1 warp)", reps=2)
    sys.exit(0)
time_plan([10000], [mk(6)], "1 unit T=10000 N=6 (1 warp)")
time_plan([10000], [mk(9)], "1 unit T=10000 N=9 (1 warp)")
time_plan([10000], [mk(12)], "1 unit T=10000 N=12 (2 warps)")
time_plan([2000], [mk(6)], "1 unit T=2000 N=6")
time_plan([10000] * 148, [mk(6) for _ in range(148)], "148 units T=10000 N=6")
time_plan([10000] * 592, [mk(6) for _ in range(592)], "592 units T=10000 N=6")
time_plan([10000] * 2368, [mk(6) for _ in range(2368)], "2368 units T=10000 N=6")
time_plan([10000] * 2368, [mk(12) for _ in range(2368)], "2368 units T=10000 N=12")
time_plan([2000] * 2368, [mk(6) for _ in range(2368)], "2368 units T=2000 N=6")
T, trs, _ = bench.make_split(0)
time_plan(T, trs, "c2 split")
