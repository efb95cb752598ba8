"""GPU scratch tool: where do one-hot operands land in mucon_wgrad_tf32's output?  python scripts/probe_wgrad.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mucon_b200 import train
from mucon_b200.temporal import BackbonePlan
dev = torch.device("cuda:0")
R = 128
plan = BackbonePlan([R], 0, dev)

def run(dY, X):
    dW = torch.zeros(128, 128, device=dev); db = torch.zeros(128, device=dev)
    train.wgrad_rows(dY, X, plan, 0, (0,), (0,), (0,), 128, dW, db)
    torch.cuda.synchronize()
    return dW.cpu().numpy(), db.cpu().numpy()

def nz(a, k=6):
    idx = np.argwhere(a != 0)
    return [(int(i), int(j), float(a[i, j])) for i, j in idx[:k]], len(idx)

for co0, ci0 in [(0, 0), (1, 0), (0, 1), (5, 9), (33, 2), (2, 33), (100, 77)]:
    dY = torch.zeros(R, 128, device=dev); X = torch.zeros(R, 128, device=dev)
    dY[:, co0] = 1; X[:, ci0] = 1
    dW, db = run(dY, X)
    print("cols", (co0, ci0), "->", nz(dW), "db nz", np.argwhere(db != 0).ravel()[:4], db.max())
for t0 in [0, 1, 7, 8, 31, 32, 100]:
    dY = torch.zeros(R, 128, device=dev); X = torch.zeros(R, 128, device=dev)
    dY[t0, 3] = 1; X[t0, 4] = 1
    dW, db = run(dY, X)
    print("row", t0, "->", nz(dW))
for t0, t1 in [(0, 1), (0, 8), (3, 5), (9, 41)]:
    dY = torch.zeros(R, 128, device=dev); X = torch.zeros(R, 128, device=dev)
    dY[t0, 3] = 1; X[t1, 4] = 1
    dW, db = run(dY, X)
    print("rows", (t0, t1), "->", nz(dW))
torch.manual_seed(0)
dY = torch.randn(R, 128, device=dev); X = torch.randn(R, 128, device=dev)
dW, db = run(dY, X)
want = (dY.double().t() @ X.double()).cpu().numpy()
print("random: |dW| max", np.abs(dW).max(), "want max", np.abs(want).max(), "err", np.abs(dW - want).max(),
      "errT", np.abs(dW.T - want).max(), "corr", np.corrcoef(dW.ravel(), want.ravel())[0, 1])
