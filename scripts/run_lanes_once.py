import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
dev = torch.device("cuda:0")
T, trs, means = bench.make_split(0)
logp = bench.device_logp(T, trs, 0, dev)
eng = ViterbiEngine(dev)
plan = AlignPlan(T, [[t.tolist()] for t in trs], 48, device=dev, len_params=poisson_params(means))
for _ in range(2):
    eng.run(plan, logp, seg0_f32=True, mode=sys.argv[1] if len(sys.argv) > 1 else "lanes")
torch.cuda.synchronize()
