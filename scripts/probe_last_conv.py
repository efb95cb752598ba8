"""GPU scratch tool: last_conv (conv_gemm_kernel, one tap) on the c2 split's deepest level, timed alone."""
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
level = max(k for k in range(8) if k < len(plan.rows))
for lvl in sorted({level, len(plan.rows) - 1}):
    rows = plan.rows[lvl]
    for kind in ("randn", "relu(randn)", "zeros"):
        x = torch.randn(rows, 128, device=dev)
        if kind == "relu(randn)":
            x = x.relu()
        if kind == "zeros":
            x.zero_()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for r in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = temporal.conv_gemm_rows(x, w["last_k"], w["last_b"], plan, lvl)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[2:])
        print(f"level {lvl} rows {rows} input {kind:12s}: {ts[len(ts) // 2]:7.1f} us  (min {ts[0]:.1f})")
