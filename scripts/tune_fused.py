"""GPU scratch tool: times the fused alignment kernel on c2 for several ring configurations."""
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
eng.run(plan, logp, seg0_f32=True, mode="split")
torch.cuda.synchronize()
ref = eng.fetch(plan)
frames = int(T.sum())


def timeit(write_bs, n=20):
    for _ in range(3):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=write_bs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=write_bs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cfgs = []
for st in (2, 3, 4, 5):
    for slab in (5760, 11520, 17280, 23040, 34560):
        for ring in (2, 4):
            cfgs.append((st, slab, ring))
for st, slab, ring in cfgs:
    os.environ["MUCON_FUSED_STAGES"] = str(st)
    os.environ["MUCON_FUSED_SLAB_BYTES"] = str(slab)
    os.environ["MUCON_FUSED_RING"] = str(ring)
    try:
        ms = timeit(False)
        out = eng.fetch(plan)
        ok = np.array_equal(out["labels"], ref["labels"]) and np.array_equal(out["score"], ref["score"])
        print(f"stages={st} slab={slab:6d} ring={ring}  {ms*1e3:7.1f} us  {frames/ms/1e6:7.2f} Gframes/s  exact={ok}", flush=True)
    except Exception as e:
        print(st, slab, ring, "FAILED", str(e)[:80], flush=True)
