"""GPU developer tool: per-role timeline of the bf16 WaveNet layer kernel (clock64 stamps of CTA 0).

    MUCON_LAYER_TRACE=1 python -m mucon_b200.build && python scripts/trace_layer16.py [dil]
    python -m mucon_b200.build          # back to the product build afterwards
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import _lib, temporal  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = temporal.MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
plan = m.plan(T)
LEVEL = int(os.environ.get("LEVEL", "0"))
x = torch.randn(plan.rows[LEVEL], 128, device=dev).to(torch.bfloat16)
w = m.ft._weights()
lib = _lib.lib()
if not hasattr(lib, "mucon_debug_layer_trace"):
    sys.exit("build with MUCON_LAYER_TRACE=1 first")
POOL = os.environ.get("POOL", "0") == "1"
for dil in [int(a) for a in sys.argv[1:]] or [1, 64]:
    wdk, w1k = w["layers_k16"][0]
    for _ in range(3):
        temporal.wavenet_layer_bf16_rows(x, wdk, w["layers_bias_h"][0][0], w1k, w["layers_bias_h"][0][1], plan, LEVEL, dil, POOL, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    temporal.wavenet_layer_bf16_rows(x, wdk, w["layers_bias_h"][0][0], w1k, w["layers_bias_h"][0][1], plan, LEVEL, dil, POOL, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"launch: {e0.elapsed_time(e1) * 1e3:.1f} us, {plan.ltiles[(LEVEL, 'same')][1]} tiles")
    buf = np.zeros((32, 128), dtype=np.int64)
    assert lib.mucon_debug_layer_trace(buf.ctypes.data_as(C.c_void_p)) == 0
    lo, hi = (10, 100) if LEVEL == 0 else (3, 22)
    d = lambda a, b, sa=0, sb=0: float(np.median(buf[a, lo + sa:hi + sa] - buf[b, lo + sb:hi + sb]))
    print(f"--- dil {dil}: cycles, medians over tiles {lo}..{hi - 1} of CTA 0")
    print("tile period (epilogue 2 done)        ", float(np.median(np.diff(buf[9, lo:hi]))))
    print("producer: stage free -> slab landed  ", d(1, 0))
    print("slab landed -> GEMM 1 may start      ", d(2, 1))
    print("GEMM 1: wait for slab/acc            ", d(2, 12))
    print("GEMM 1 issue span                    ", d(3, 2))
    print("GEMM 1 issued -> epilogue 1 starts   ", d(6, 3))
    print("epilogue 1                           ", d(7, 6))
    print("  E1: residual -> acc2 issued          ", d(10, 6))
    print("  E1: acc1 loaded                      ", d(11, 10))
    print("  E1: partner barrier                  ", d(14, 11))
    print("  E1: bias/relu/pack + st + wait       ", d(7, 14))
    print("  E2: first TMEM load                  ", d(15, 8))
    print("  E2: first pack + staging             ", d(16, 15))
    print("  E2: second load + pack + staging     ", d(17, 16))
    print("  E2: proxy fence                      ", d(18, 17))
    print("  E2: barrier                          ", d(19, 18))
    print("  E2: TMA store issue                  ", d(9, 19))
    print("  E1a: slab rows loaded + unpacked     ", d(20, 6))
    print("  E1a: first STTM issued               ", d(21, 20))
    print("GEMM 2: wait for Y                   ", d(4, 13))
    print("epilogue 1 done -> GEMM 2 may start  ", d(4, 7))
    print("GEMM 2 issue span                    ", d(5, 4))
    print("GEMM 2 issued -> epilogue 2 starts   ", d(8, 5))
    print("epilogue 2                           ", d(9, 8))
    print("GEMM 2 issued(i) -> stage free(i+2)  ", d(0, 5, 2, 0))
    print("CTA 0 span: first slab request -> last epilogue 2:", int(buf[9].max() - buf[0, 0]), "cycles")
    print("first tiles, relative to producer tile 0:")
    for ev in (0, 1, 12, 2, 3, 6, 7, 13, 4, 5, 8, 9):
        print(f"  ev{ev:2d}", (buf[ev, :6] - buf[0, 0]).tolist())
