"""Multi-GPU plumbing: one process per GPU, videos sharded across ranks, ONE collective at the end.

(video, candidate) units are independent (the reference is strictly batch-size 1,
src/core/datasets/general_dataset.py:169-172, src/mucon/evaluators.py:260), so the data path has
no exchange step: each rank aligns its own videos and only the per-video alignments (segment
lengths in blocks -- the run-length form of the frame labels) and path scores are gathered with
NCCL over NVLink.  All candidates of a video stay on one rank because they share its block-score
table.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_videos(T, n_cands, world_size):
    """Greedy longest-first balancing of work (T_v * candidates_v) over ranks.

    Returns a list of index arrays (ascending within a rank), one per rank."""
    T = np.asarray(T, dtype=np.int64)
    w = T * np.asarray(n_cands, dtype=np.int64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    bins = [[] for _ in range(world_size)]
    for v in order:
        r = int(np.argmin(load))
        bins[r].append(int(v))
        load[r] += w[v]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def gather_payload(plan, recv=None, group=None):
    """The one collective of the data path: all_gather of every rank's AlignPlan.payload
    ([score f64 x U | seg_blocks i32 x sum N | padding], written in place by the kernels; all ranks
    must have built their plans with the same payload_capacity).  Returns [world, capacity] uint8."""
    world = dist.get_world_size(group)
    n = plan.payload.numel()
    if recv is None:
        recv = torch.empty(world * n, dtype=torch.uint8, device=plan.payload.device)
    dist.all_gather_into_tensor(recv, plan.payload, group=group)
    return recv.view(world, n)


class PipelinedGather:
    """gather_payload for a stream of batches: the collective of batch i runs on the communicator's
    stream while batch i+1 is being aligned.  Batches alternate between `depth` (plan, receive buffer)
    slots; before a slot's plan is run again the gather that still reads its payload is waited for
    (a stream-side wait, the host does not block).

        pg = PipelinedGather([plan_a, plan_b])
        for i, logp in enumerate(batches):
            plan = pg.acquire(i)               # safe to overwrite this slot's payload now
            engine.run(plan, logp, ...)
            pg.gather(i)                       # async all_gather of plan.payload
        rows = pg.result(i)                    # [world, capacity] uint8, waits for batch i's gather
    """

    def __init__(self, plans, group=None):
        self.plans, self.group = list(plans), group
        world = dist.get_world_size(group)
        n = self.plans[0].payload.numel()
        if any(p.payload.numel() != n for p in self.plans):
            raise ValueError("all plans need the same payload_capacity")
        self.recv = [torch.empty(world * n, dtype=torch.uint8, device=p.payload.device) for p in self.plans]
        self.work = [None] * len(self.plans)
        self.world, self.n = world, n

    def acquire(self, i):
        s = i % len(self.plans)
        if self.work[s] is not None:
            self.work[s].wait()
            self.work[s] = None
        return self.plans[s]

    def gather(self, i):
        s = i % len(self.plans)
        self.work[s] = dist.all_gather_into_tensor(self.recv[s], self.plans[s].payload, group=self.group, async_op=True)

    def result(self, i):
        s = i % len(self.plans)
        if self.work[s] is not None:
            self.work[s].wait()
            self.work[s] = None
        return self.recv[s].view(self.world, self.n)

    def drain(self):
        for s in range(len(self.plans)):
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None


def unpack_payload(row, U, n_pos):
    """(scores [U] float64, seg_blocks [n_pos] int32) views of one rank's gathered payload row."""
    return row[:8 * U].view(torch.float64), row[8 * U:8 * U + 4 * n_pos].view(torch.int32)


def gather_alignments(score, seg_blocks, tr_off, max_units, max_positions, group=None):
    """all_gather of per-unit scores [U] (float64) and segment lengths [sum N] (int32) plus the
    transcript offsets, padded to the given maxima.  Returns lists indexed by rank of
    (score, seg_blocks, tr_off) trimmed back to each rank's true sizes."""
    world = dist.get_world_size(group)
    dev = score.device
    U, P = score.shape[0], seg_blocks.shape[0]
    sc = torch.zeros(max_units, dtype=torch.float64, device=dev)
    sc[:U] = score
    ints = torch.zeros(2 + max_positions + max_units + 1, dtype=torch.int32, device=dev)
    ints[0], ints[1] = U, P
    ints[2:2 + P] = seg_blocks
    ints[2 + max_positions:2 + max_positions + U + 1] = torch.as_tensor(tr_off, dtype=torch.int32, device=dev)
    all_sc = torch.empty(world * max_units, dtype=torch.float64, device=dev)
    all_int = torch.empty(world * ints.shape[0], dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_sc, sc, group=group)
    dist.all_gather_into_tensor(all_int, ints, group=group)
    return all_sc.view(world, max_units), all_int.view(world, -1)


def unpack_gathered(all_sc, all_int, max_positions):
    out = []
    for r in range(all_sc.shape[0]):
        row = all_int[r].cpu().numpy()
        U, P = int(row[0]), int(row[1])
        out.append(dict(score=all_sc[r, :U].cpu().numpy(), seg_blocks=row[2:2 + P],
                        tr_off=row[2 + max_positions:2 + max_positions + U + 1]))
    return out
