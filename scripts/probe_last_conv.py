"""GPU scratch tool: the one-tap TF32 conv_gemm_kernel (last_conv of the TF32 path) per level of the c2 split and on
synthetic plans of full / partly filled tiles, timed alone with the L2 flushed: microseconds per 128-row tile."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import temporal  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = temporal.MuConBackbone().eval().to(dev)
w = m.ft._weights()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(plan, lvl, name):
    rows = plan.rows[lvl]
    x = torch.randn(rows, 128, device=dev)
    ts = []
    for r in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        temporal.conv_gemm_rows(x, w["last_k"], w["last_b"], plan, lvl)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    t = sorted(ts[2:])[2]
    nt = plan.n_tiles[lvl]
    print(f"{name:34s} rows {rows:8d} tiles {nt:6d}: {t:8.1f} us = {t / max(1, -(-nt // 148)):6.2f} us per tile and CTA")


T, trs, _ = bench.make_split(0)
plan = m.plan(T)
for lvl in range(len(plan.rows)):
    run(plan, lvl, f"c2 level {lvl}")
for Tv, V in ((128, 2720), (140, 1712), (256, 1360), (64, 2720), (1280, 272), (12800, 27)):
    p = temporal.BackbonePlan(np.full(V, Tv), 0, dev)
    run(p, 0, f"{V} videos of {Tv} rows")
