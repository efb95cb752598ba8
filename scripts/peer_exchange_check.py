"""Multi-GPU check of dist.PeerExchange (run under torchrun, one rank per GPU):
every rank aligns its own small batch; after the fence each rank must hold every rank's payload, equal to what an
NCCL all_gather of the payloads returns.  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_exchange_check.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mucon_b200 import dist as mdist
from mucon_b200.length_model import poisson_params
from mucon_b200.viterbi import AlignPlan, ViterbiEngine
from tests import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, Cn = 40, 48
T, trs = synth.breakfast_split(seed=10 + rank, V=V, C=Cn, fs=30, J=66)
rng = np.random.default_rng(rank)
means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, Cn, int(t)) for tr, t in zip(trs, T)])
cap = 8 * V + 4 * V * 12
plans = [AlignPlan(T, [[tr.tolist()] for tr in trs], Cn, device=dev, len_params=poisson_params(means), labels="best",
                   payload_capacity=cap) for _ in range(2)]
logp = torch.log_softmax(torch.randn(int(T.sum()), Cn, device=dev, generator=torch.Generator(dev).manual_seed(rank)), 1)
eng = ViterbiEngine(dev)
px = mdist.PeerExchange(plans)
ok = True
for i in range(4):
    eng.run(px.plan(i), logp, seg0_f32=True, write_bs=False)
    px.fence()
    want = mdist.gather_payload(px.plan(i))
    torch.cuda.synchronize()
    ok &= bool(torch.equal(px.result(i), want))
    ok &= bool(px.result(i)[rank].view(torch.uint8)[:8 * V].view(torch.float64).isfinite().all())
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
px.close()
if rank == 0:
    print(json.dumps({"peer_exchange_ok": bool(flag.item()), "world": world}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
