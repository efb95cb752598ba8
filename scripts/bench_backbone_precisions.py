"""GPU scratch tool: per-kernel timing of the backbone forward on the c2 split for each precision of the layers."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import temporal  # noqa: E402
from mucon_b200.temporal import MuConBackbone  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=5, warm=2):
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


torch.manual_seed(0)
m = MuConBackbone().eval().to(dev)
T, trs, _ = bench.make_split(0)
nv = int(os.environ.get("NV", "1712"))
Ts = T[:nv]
plan = m.plan(Ts)
feats = torch.randn(int(Ts.sum()), 2048, device=dev).abs_() * 0.5
out = {}
for prec in ("fp16", "bf16", "tf32"):
    t_enc = timeit(lambda: m.encode_packed(feats, plan, precision=prec))
    z = m.encode_packed(feats, plan, precision=prec)
    t_cls = timeit(lambda: m.logprobs_packed(z, plan))
    w = m.ft._weights()
    t_proj = timeit(lambda: temporal.gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], True, out_dtype={"fp16": torch.float16, "bf16": torch.bfloat16, "tf32": torch.float32}[prec]))
    out[prec] = dict(encode_ms=t_enc, classifier_logsoftmax_ms=t_cls, projection_ms=t_proj)
    print(f"{prec}: encode {t_enc:.3f} ms (projection {t_proj:.3f}) classifier+logsoftmax {t_cls:.3f} ms", flush=True)
    if prec == "bf16":
        x = temporal.gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], True, out_bf16=True)
        level, last = 0, len(m.ft.stages) - 1
        for i, (wd, bd, w1, b1) in enumerate(w["layers"]):
            pooled = i in m.ft.pooling_layers
            wdk, w1k = w["layers_k16"][i]
            bd, b1 = w["layers_bias_h"][i]
            f = lambda: temporal.wavenet_layer_bf16_rows(x, wdk, bd, w1k, b1, plan, level, m.ft.stages[i], pooled,
                                                         relu_final=(i == last), out_f32=(i == last and not pooled))
            t = timeit(f)
            rows = x.shape[0]
            byts = rows * 256 + (rows // 2 if pooled else rows) * (512 if i == last else 256)
            print(f"   layer {i:2d} dil {m.ft.stages[i]:4d} rows {rows:8d} pool {int(pooled)}: {t*1e3:8.1f} us  "
                  f"{byts/t/1e6:7.1f} GB/s  {rows/128/148*1.0:6.1f} tiles/SM  {t*1e-3*1.9e9/(rows/128/148):7.0f} cyc/tile",
                  flush=True)
            out[prec][f"layer{i}_us"] = t * 1e3
            x = f()
            if pooled:
                level += 1
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/backbone_precisions.json", "w"), indent=1)
