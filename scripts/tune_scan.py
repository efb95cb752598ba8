"""GPU scratch tool: times the block-score scan kernel alone on the c2 workload for several
staging configurations (MUCON_SCAN_MODE / _STAGES / _SLAB_BYTES).  Not part of the product."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mucon_b200 import _lib  # noqa: E402
from mucon_b200.length_model import poisson_params  # noqa: E402
from mucon_b200.viterbi import AlignPlan, ViterbiEngine  # noqa: E402

dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
logp = bench.device_logp(T, trs, 0, dev)
eng = ViterbiEngine(dev)
plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means))
eng.run(plan, logp, seg0_f32=True)
torch.cuda.synchronize()
ref_bs = plan.bs.clone()
lib = _lib.lib()
p = plan.p
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
nbytes = logp.numel() * 4


def scan():
    _lib.check(lib.mucon_viterbi_blockscores(_lib.ptr(logp), 0, C.c_void_p(p["vid_off"]), C.c_void_p(p["blk_off"]),
                                             C.c_void_p(p["order_v"]), plan.V, 48, 30, _lib.ptr(plan.bs), sp), "scan")


def timeit(n=20):
    for _ in range(3):
        scan()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        scan()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


configs = [dict(MUCON_SCAN_MODE="1")]
for st in (2, 3, 4, 6, 8, 12, 16):
    for slab in (5760, 11520, 23040):
        configs.append(dict(MUCON_SCAN_MODE="2", MUCON_SCAN_STAGES=str(st), MUCON_SCAN_SLAB_BYTES=str(slab)))
for cfg in configs:
    for k in ("MUCON_SCAN_MODE", "MUCON_SCAN_STAGES", "MUCON_SCAN_SLAB_BYTES"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    try:
        ms = timeit()
        ok = torch.equal(plan.bs, ref_bs)
        print(f"{cfg}  {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  exact={ok}", flush=True)
    except Exception as e:
        print(cfg, "FAILED", e, flush=True)
