"""GPU scratch tool: per-layer timings of the 16-bit WaveNet layer kernel on the c2 split (A/B of two builds:
swap mucon_b200/libmucon_b200.so between runs on the same box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import temporal  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = temporal.MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
plan = m.plan(T)
w = m.ft._weights()
wdk, w1k = w["layers_k16"][0]
bd, b1 = w["layers_bias_h"][0]
x = torch.randn(int(T.sum()), 128, device=dev).to(wdk.dtype)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for dil, pool in ((1, False), (2, False), (4, True), (32, False), (64, False), (512, False)):
    ts = []
    for r in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        temporal.wavenet_layer_bf16_rows(x, wdk, bd, w1k, b1, plan, 0, dil, pool, False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    print(f"level 0 dil {dil:4d} pool {int(pool)}: {ts[len(ts) // 2] * 1e3:8.1f} us")

# cost of the residual MMAs: the same launches without the skip connection
for dil, pool in ((1, False), (4, True), (64, False)):
    for res in (True, False):
        ts = []
        for r in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            temporal.wavenet_layer_bf16_rows(x, wdk, bd, w1k, b1, plan, 0, dil, pool, False, residual=res)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        print(f"level 0 dil {dil:4d} pool {int(pool)} residual {int(res)}: {ts[len(ts) // 2] * 1e3:8.1f} us")
