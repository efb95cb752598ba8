"""GPU scratch tool: eager c5 training steps for an ncu launch list.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/prof_train.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200 import train
from mucon_b200.temporal import MuConBackbone
dev = torch.device("cuda:0")
T_all, _, _ = bench.make_split(0)
rng = np.random.default_rng(5)
Ts = [int(t) for t in T_all[:32]]
Ns = [int(rng.integers(2, 13)) for _ in Ts]
torch.manual_seed(0)
m = MuConBackbone().to(dev).train()
opt = torch.optim.SGD(m.parameters(), lr=1e-3)
ts = train.TrainStep(m, Ts, Ns, optimizer=opt, graph=False)
ts.feats.copy_(torch.randn(ts.feats.shape, device=dev).abs() * 0.5)
ts.transcripts.copy_(torch.from_numpy(np.concatenate([rng.integers(0, 48, n) for n in Ns])).to(dev))
for _ in range(int(os.environ.get("STEPS", "2"))):
    ts.run()
torch.cuda.synchronize()
print("loss", ts.loss.item())
