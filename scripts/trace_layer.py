"""GPU developer tool: per-role timeline of the WaveNet layer kernel (clock64 stamps of CTA 0).

    MUCON_LAYER_TRACE=1 python -m mucon_b200.build && MUCON_LAYER_SLAB=0 python scripts/trace_layer.py
    python -m mucon_b200.build          # back to the product build afterwards

MUCON_LAYER_SLAB=0 routes the small dilations through wavenet_layer_kernel (the instrumented one)."""
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
x = torch.randn(int(T.sum()), 128, device=dev)
w = m.ft._weights()
wdk, w1k = w["layers_k"][0]
for _ in range(3):
    temporal.wavenet_layer_rows(x, wdk, w["layers"][0][1], w1k, w["layers"][0][3], plan, 0, 1, False, False, pair=False)
torch.cuda.synchronize()
lib = _lib.lib()
if not hasattr(lib, "mucon_debug_layer_trace"):
    sys.exit("build with MUCON_LAYER_TRACE=1 first")
buf = np.zeros((32, 128), dtype=np.int64)
assert lib.mucon_debug_layer_trace(buf.ctypes.data_as(C.c_void_p)) == 0
d = lambda a, b: float(np.median(buf[a, 10:100] - buf[b, 10:100]))
print("cycles, medians over tiles 10..99 of CTA 0")
print("tile period                          ", float(np.median(np.diff(buf[10, 10:100]))))
print("GEMM 1 issue span                    ", d(3, 2))
print("GEMM 2 issue span                    ", d(5, 4))
print("epilogue 1 (a1full -> Y ready)       ", d(7, 6))
print("epilogue 2: TMEM -> staging          ", d(9, 8))
print("epilogue 2: residual + stores        ", d(11, 9))
print("epilogue 1 start - GEMM 1 issued     ", d(6, 3))
print("GEMM 2 start - epilogue 1 done       ", d(4, 7))
print("next GEMM 1 start - GEMM 2 issued    ", float(np.median(buf[2, 11:100] - buf[5, 10:99])))
print("next epilogue 1 start - tile done    ", float(np.median(buf[6, 11:100] - buf[10, 10:99])))
