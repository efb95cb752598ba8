"""GPU scratch tool: WaveNet layers with and without the cluster-pair weight multicast (c2 batch)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200 import temporal
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = temporal.MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
plan = m.plan(T)
feats = torch.randn(int(T.sum()), 2048, device=dev).abs_() * 0.5
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ref = None
for pairs in (False, True, False, True):
    temporal.LAYER_PAIRS = pairs
    ms = timeit(lambda: m.encode_packed(feats, plan))
    z = m.encode_packed(feats, plan)
    if ref is None: ref = z
    print(f"pairs={pairs}: encode {ms:7.3f} ms  max|diff| vs first = {(z-ref).abs().max().item():.3e}", flush=True)
w = m.ft._weights()
x = temporal.gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], True)
wdk, w1k = w["layers_k"][0]
for pairs in (False, True):
    ms = timeit(lambda: temporal.wavenet_layer_rows(x, wdk, w["layers"][0][1], w1k, w["layers"][0][3], plan, 0, 1, False, False, pair=pairs), n=10)
    print(f"level-0 layer pairs={pairs}: {ms*1e3:8.1f} us", flush=True)
