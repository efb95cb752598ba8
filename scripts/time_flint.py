import os, sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from mucon_b200.loss import _flint_meta, flint_evidence
dev = torch.device("cuda:0")
T, trs, _ = bench.make_split(0)
Ms = [len(t) for t in trs]; Tl = [int(t) for t in T]
rng = np.random.default_rng(1000)
L = torch.from_numpy(np.concatenate([float(t) * rng.dirichlet(3 * np.ones(m)) for t, m in zip(T, Ms)]).astype(np.float32)).to(dev)
seg = torch.randn(int(T.sum()), 48, device=dev)
meta = _flint_meta(Ms, Tl, dev)
for _ in range(3): flint_evidence(L, seg, Ms, Tl, meta=meta)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): flint_evidence(L, seg, Ms, Tl, meta=meta)
b.record(); torch.cuda.synchronize()
print(os.environ.get("MUCON_FLINT_PARTS"), "ms", a.elapsed_time(b) / 20)
