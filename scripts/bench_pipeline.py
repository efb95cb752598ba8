"""GPU: backbone inference with the projection / layer kernels side by side on disjoint SMs (infer_pooled_pipelined)
against the plain sequence, c2 split."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.temporal import MuConBackbone
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
plan = m.plan(T)
feats = torch.randn(int(T.sum()), 2048, device=dev).abs_() * 0.5
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
base = timed(lambda: m.infer_pooled_packed(feats, plan))
ref, _ = m.infer_pooled_packed(feats, plan)
print(f"sequential: {base:.3f} ms")
for chunks in (3, 4, 6):
    for p in (40, 48, 56, 64, 72):
        ms = timed(lambda: m.infer_pooled_pipelined(feats, T, n_chunks=chunks, proj_sms=p))
        tb, off = m.infer_pooled_pipelined(feats, T, n_chunks=chunks, proj_sms=p)
        torch.cuda.synchronize()
        print(f"chunks {chunks} proj_sms {p}: {ms:.3f} ms  same={bool(torch.equal(tb, ref))}")
