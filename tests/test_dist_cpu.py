"""CPU: the multi-GPU plumbing on gloo with world_size 2 (sharding + the single gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mucon_b200 import dist as mdist


def test_shard_videos_balances_and_partitions():
    rng = np.random.default_rng(0)
    T = rng.integers(300, 10000, 500)
    nc = rng.integers(1, 5, 500)
    for world in (1, 2, 4, 8):
        parts = mdist.shard_videos(T, nc, world)
        allv = np.sort(np.concatenate(parts))
        assert np.array_equal(allv, np.arange(500))  # every video on exactly one rank
        load = np.array([(T[p] * nc[p]).sum() for p in parts])
        assert load.max() <= 1.02 * load.mean() + (T * nc).max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    U = 5 + 3 * rank
    N = rng.integers(1, 7, U)
    tr_off = np.concatenate([[0], np.cumsum(N)]).astype(np.int32)
    score = torch.from_numpy(rng.standard_normal(U))
    seg = torch.from_numpy(rng.integers(1, 66, int(tr_off[-1])).astype(np.int32))
    all_sc, all_int = mdist.gather_alignments(score, seg, tr_off, max_units=16, max_positions=128)
    got = mdist.unpack_gathered(all_sc, all_int, 128)
    ok = True
    for r in range(world):
        rr = np.random.default_rng(100 + r)
        Ur = 5 + 3 * r
        Nr = rr.integers(1, 7, Ur)
        offr = np.concatenate([[0], np.cumsum(Nr)]).astype(np.int32)
        scr = rr.standard_normal(Ur)
        segr = rr.integers(1, 66, int(offr[-1])).astype(np.int32)
        ok &= np.array_equal(got[r]["score"], scr) and np.array_equal(got[r]["seg_blocks"], segr)
        ok &= np.array_equal(got[r]["tr_off"], offr)
    # the zero-copy variant: every rank's plan payload gathered as bytes
    class P:
        pass
    plan = P()
    cap = 8 * 16 + 4 * 128
    plan.payload = torch.zeros(cap, dtype=torch.uint8)
    plan.payload[:8 * U].view(torch.float64).copy_(score)
    plan.payload[8 * U:8 * U + 4 * int(tr_off[-1])].view(torch.int32).copy_(seg)
    rows = mdist.gather_payload(plan)
    for r in range(world):
        rr = np.random.default_rng(100 + r)
        Ur = 5 + 3 * r
        Nr = rr.integers(1, 7, Ur)
        npos = int(Nr.sum())
        scr = rr.standard_normal(Ur)
        segr = rr.integers(1, 66, npos).astype(np.int32)
        sc_v, sg_v = mdist.unpack_payload(rows[r], Ur, npos)
        ok &= np.array_equal(sc_v.numpy(), scr) and np.array_equal(sg_v.numpy(), segr)
    # pipelined gather: two slots, batch i's collective overlaps batch i+1's work
    plans = []
    for _ in range(2):
        q = P()
        q.payload = torch.zeros(64, dtype=torch.uint8)
        plans.append(q)
    pg = mdist.PipelinedGather(plans)
    for i in range(5):
        q = pg.acquire(i)
        q.payload.fill_(10 * i + rank)           # "the alignment of batch i"
        pg.gather(i)
        if i >= 1:                                # batch i-1's result is complete and untouched by batch i
            rows = pg.result(i - 1)
            ok &= all(int(rows[r][0]) == 10 * (i - 1) + r and int(rows[r][-1]) == 10 * (i - 1) + r for r in range(world))
    rows = pg.result(4)
    ok &= all(int(rows[r][7]) == 40 + r for r in range(world))
    pg.drain()
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_alignments_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_peer_deltas_layout():
    """byte offsets from a payload buffer to this rank's slot in every rank's receive buffer ([slot][rank][cap])"""
    from mucon_b200.dist import peer_deltas
    d = peer_deltas(payload_ptr=1000, peer_recv_ptrs=[50000, 90000], rank=1, world=2, slot=1, capacity=256)
    assert d == [50000 + (1 * 2 + 1) * 256 - 1000, 90000 + (1 * 2 + 1) * 256 - 1000]
