import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
logp = bench.device_logp(T, trs, 0, dev)
eng = ViterbiEngine(dev)
plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means))
eng.run(plan, logp, seg0_f32=True, mode="split"); torch.cuda.synchronize()
ref = eng.fetch(plan, want_bp=True)
frames = int(T.sum())
for cpt in (1, 2):
    for st, slab, ring in ((2, 17280, 4), (3, 17280, 4), (2, 23040, 2), (2, 11520, 4), (3, 11520, 4), (2, 23040, 4), (3, 23040, 2)):
        os.environ.update(MUCON_FUSED_CPT=str(cpt), MUCON_FUSED_STAGES=str(st), MUCON_FUSED_SLAB_BYTES=str(slab), MUCON_FUSED_RING=str(ring))
        for _ in range(3):
            eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=True)
        torch.cuda.synchronize()
        out = eng.fetch(plan, want_bp=True)
        ok = all(np.array_equal(out[k], ref[k]) for k in ("labels", "score", "bp", "bs", "seg_blocks"))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"cpt={cpt} stages={st} slab={slab:6d} ring={ring}  {ms*1e3:7.1f} us  {frames/ms/1e6:7.2f} Gframes/s exact={ok}", flush=True)
