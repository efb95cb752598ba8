"""GPU scratch tool: the s-head on the c2 split (teacher-forced), for ncu launch lists."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.shead import SHead
dev = torch.device("cuda:0")
T, trs, _ = bench.make_split(0)
Tz = T // 16
off_h = np.concatenate([[0], np.cumsum(Tz)]).astype(np.int64)
off = torch.from_numpy(off_h).to(dev)
torch.manual_seed(0)
sh = SHead().eval().to(dev)
z = torch.randn(int(Tz.sum()), 128, device=dev).relu()
tf = [np.concatenate([[49], tr]).astype(np.int32) for tr in trs]
for _ in range(2):
    out = sh.forward_packed(z, off, off_h, tf, teacher_forcing=True)
torch.cuda.synchronize()
