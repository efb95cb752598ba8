"""GPU scratch tool: two backbone forwards on the c2 split through the current fast path (for ncu launch lists).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_backbone.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.temporal import MuConBackbone
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
nv = int(os.environ.get("NV", "1712"))
Ts = T[:nv]
plan = m.plan(Ts)
feats = torch.randn(int(Ts.sum()), 2048, device=dev).abs_() * 0.5
for _ in range(2):
    table, off = m.infer_pooled_packed(feats, plan)
torch.cuda.synchronize()
