"""GPU scratch tool: mask kernels on the c2 split, kernel-level timings."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200 import masks as mm
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
Ms = [len(t) for t in trs]
rng = np.random.default_rng(0)
Lh = np.concatenate([float(t) * rng.dirichlet(3 * np.ones(m)) for t, m in zip(T, Ms)]).astype(np.float32)
L = torch.from_numpy(Lh).to(dev).requires_grad_(True)
Ts = [int(t) for t in T]
nbytes = sum(4 * t * m for t, m in zip(Ts, Ms))
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(3):
    out, off = mm.create_masks_batch(Ts, L, Ms)
torch.cuda.synchronize()
a, b = ev(), ev()
t0 = time.perf_counter(); a.record()
for _ in range(10):
    out, off = mm.create_masks_batch(Ts, L, Ms)
b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"fwd (python entry): device {a.elapsed_time(b)/10*1e3:8.1f} us  host {1e5*(t1-t0):8.1f} us  {nbytes/1e6:.1f} MB", flush=True)
go = torch.ones_like(out)
for _ in range(2):
    L.grad = None; out.backward(go, retain_graph=True)
torch.cuda.synchronize()
a, b = ev(), ev()
t0 = time.perf_counter(); a.record()
for _ in range(5):
    L.grad = None; out.backward(go, retain_graph=True)
b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"bwd (autograd): device {a.elapsed_time(b)/5*1e3:8.1f} us  host {2e5*(t1-t0):8.1f} us", flush=True)
