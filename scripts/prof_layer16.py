"""GPU scratch tool: a few launches of the bf16 WaveNet layer kernel on the c2 split, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:wavenet_layer_bf16 -s 2 -c 1 -o gpurun_out/layer16 python scripts/prof_layer16.py [dil] [pool]"""
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
x = torch.randn(int(T.sum()), 128, device=dev).to(torch.bfloat16)
w = m.ft._weights()
dil = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pool = len(sys.argv) > 2 and sys.argv[2] == "1"
wdk, w1k = w["layers_k16"][0]
bd, b1 = w["layers_bias_h"][0]
for _ in range(4):
    temporal.wavenet_layer_bf16_rows(x, wdk, bd, w1k, b1, plan, 0, dil, pool, False)
torch.cuda.synchronize()
