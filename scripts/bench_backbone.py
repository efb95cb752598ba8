"""GPU scratch tool: timings of the backbone forward pieces (projection GEMM on tcgen05, fp32 layers)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200.temporal import MuConBackbone, gemm_tf32_bias_act  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for M in (2000, 131072, 1048576, 3853222):
    A = torch.randn(M, 2048, device=dev).abs_()
    W = torch.randn(128, 2048, device=dev) / 45
    b = torch.randn(128, device=dev)
    ms = timeit(lambda: gemm_tf32_bias_act(A, W, b, True))
    byts = M * 2048 * 4 + M * 128 * 4 + 128 * 2048 * 4
    fl = 2.0 * M * 2048 * 128
    print(f"proj GEMM M={M:8d}: {ms*1e3:9.1f} us  {byts/ms/1e6:8.1f} GB/s ({byts/ms/1e6/peaks['hbm_gbs']*100:5.1f}% of measured HBM)  "
          f"{fl/ms/1e9:8.1f} TFLOP/s tf32   {M/ms/1e6:7.3f} Gframes/s", flush=True)
    ms2 = timeit(lambda: torch.relu(torch.nn.functional.linear(A, W, b)))
    print(f"   torch fp32 linear+relu (cuBLAS): {ms2*1e3:9.1f} us", flush=True)
    del A

torch.manual_seed(0)
m = MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
for nv in (1, 64, 512, 1712):
    Ts = T[:nv] if nv > 1 else np.array([2000])
    plan = m.plan(Ts)
    feats = torch.randn(int(Ts.sum()), 2048, device=dev).abs_() * 0.5
    t_all = timeit(lambda: m.logprobs_packed(m.encode_packed(feats, plan), plan), n=5, warm=2)
    t_proj = timeit(lambda: gemm_tf32_bias_act(feats, m.ft._weights()["first_w"], m.ft._weights()["first_b"], True), n=5, warm=2)
    print(f"backbone fwd {nv:5d} videos {int(Ts.sum()):8d} frames: total {t_all:8.3f} ms (projection {t_proj:7.3f} ms)  "
          f"{Ts.sum()/t_all/1e6:7.3f} Gframes/s", flush=True)
    del feats
