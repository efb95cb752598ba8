import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
eng = ViterbiEngine(dev)
order = np.argsort(-T)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
idx = order[:n]
Ts = T[idx]; trl = [trs[i] for i in idx]; ms_ = means[idx]
logp = bench.device_logp(Ts, trl, 0, dev)
plan = AlignPlan(Ts, [[t.tolist()] for t in trl], 48, device=dev, len_params=poisson_params(ms_))
for _ in range(3):
    eng.run(plan, logp, seg0_f32=True, mode="fused", write_bs=False)
torch.cuda.synchronize()
