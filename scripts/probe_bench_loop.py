"""GPU scratch tool: why does bench.py's loop see a different step time than a plain loop?"""
import os, sys
import numpy as np
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
import time
for per_step_events in (False, True):
    for steps in (20, 200):
        for _ in range(3):
            eng.run(plan, logp, seg0_f32=True, write_bs=False)
        torch.cuda.synchronize()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(steps):
            if per_step_events: ev[i][0].record()
            eng.run(plan, logp, seg0_f32=True, write_bs=False)
            if per_step_events: ev[i][1].record()
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        per = np.mean([e[0].elapsed_time(e[1]) for e in ev]) * 1e3 if per_step_events else float("nan")
        print(f"events={per_step_events} steps={steps}: {a.elapsed_time(b)/steps*1e3:7.1f} us/step (device), per-step events {per:7.1f} us, host submit {1e6*(t1-t0)/steps:7.1f} us/step", flush=True)
