"""GPU scratch tool: the fused flint evidence forward on the c2 split, both kernels, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:flint_fwd -o gpurun_out/flint python scripts/prof_flint.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200 import _lib
from mucon_b200.loss import _flint_meta, flint_evidence
dev = torch.device("cuda:0")
T, trs, _ = bench.make_split(0)
Ms = [len(t) for t in trs]; Tl = [int(t) for t in T]
rng = np.random.default_rng(1000)
L = torch.from_numpy(np.concatenate([float(t) * rng.dirichlet(3 * np.ones(m)) for t, m in zip(T, Ms)]).astype(np.float32)).to(dev)
seg = torch.randn(int(T.sum()), 48, device=dev)
meta = _flint_meta(Ms, Tl, dev)
for _ in range(3):
    E = flint_evidence(L, seg, Ms, Tl, meta=meta)          # warp-per-item kernel
E0 = torch.empty_like(E)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(3):                                          # CTA-per-row kernel
    _lib.check(_lib.lib().mucon_flint_fwd(_lib.ptr(L), _lib.ptr(meta["n_off"]), _lib.ptr(meta["T"]), _lib.ptr(meta["seg_off"]),
               _lib.ptr(meta["row_vid"]), C.c_int(meta["V"]), C.c_int(meta["n_rows"]), C.c_int(48), C.c_float(0.0), C.c_int(0),
               C.c_int(0), _lib.ptr(seg), _lib.ptr(E0), st), "flint")
torch.cuda.synchronize()
print("max diff", (E - E0).abs().max().item(), E.abs().max().item())
