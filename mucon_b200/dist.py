"""Multi-GPU plumbing: one process per GPU, videos sharded across ranks, ONE collective at the end.

(video, candidate) units are independent (the reference is strictly batch-size 1,
src/core/datasets/general_dataset.py:169-172, src/mucon/evaluators.py:260), so the data path has
no exchange step: each rank aligns its own videos and only the per-video alignments (segment
lengths in blocks -- the run-length form of the frame labels) and path scores are gathered with
NCCL over NVLink.  All candidates of a video stay on one rank because they share its block-score
table.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import ctypes as C

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

    def gather(self, i, stream=None):
        """stream: the stream the plan was run on when it is not the current one -- the collective is ordered after
        the kernels that write the payload through an event on that stream."""
        s = i % len(self.plans)
        if stream is not None and self.plans[s].payload.is_cuda and stream != torch.cuda.current_stream():
            ev = torch.cuda.Event()
            ev.record(stream)
            torch.cuda.current_stream().wait_event(ev)
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
    transcript offsets, padded to the given maxima.  Returns the raw padded tensors (all_sc [world, max_units],
    all_int [world, ...]); unpack_gathered(all_sc, all_int, max_positions) trims them to per-rank dicts."""
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


def peer_deltas(payload_ptr, peer_recv_ptrs, rank, world, slot, capacity):
    """Byte offsets from a plan's payload buffer to this rank's slot in every rank's receive buffer (own rank
    included): receive buffers are laid out [slot][source rank][capacity]."""
    return [int(base) + (slot * world + rank) * capacity - int(payload_ptr) for base in peer_recv_ptrs]


class _DevBuf:
    """torch view of raw device memory (allocated by mucon_peer_alloc) through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The multi-GPU result exchange without a collective kernel: every rank's alignment kernels store the per-video
    scores and segment lengths straight into their slot of EVERY rank's receive buffer (peer-to-peer stores over
    NVLink from the kernel epilogue, mucon_viterbi_batch.peer_delta), so an all_gather of the payloads has happened
    by the time the kernels of all ranks have finished.  Replaces the NCCL all_gather_into_tensor per step
    (SURVEY.md 8e), which cost a collective kernel + a stream wait every 150 us step.

        px = PeerExchange([plan_a, plan_b])        # collective: allocates, exchanges IPC handles, opens peers
        for i, logp in enumerate(batches):
            engine.run(px.plan(i), logp, ...)      # results land on every rank
        px.fence()                                 # all ranks' kernels done (stream sync + barrier)
        rows = px.result(i)                        # [world, capacity] uint8: rank r's payload of batch i

    Slots alternate between the plans: fence() before a slot is read, and before it is overwritten a second time
    if a consumer is still reading.  Needs one node (CUDA IPC) and the nccl (or any) process group for the handle
    exchange and the fences."""

    def __init__(self, plans, group=None):
        from . import _lib
        self.lib, self.group = _lib.lib(), group
        self.plans = list(plans)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise _lib.MuconError("peer exchange covers one node (<= 8 ranks)")
        self.cap = self.plans[0].payload.numel()
        if any(p.payload.numel() != self.cap for p in self.plans):
            raise ValueError("all plans need the same payload_capacity")
        dev = self.plans[0].payload.device
        nbytes = len(self.plans) * self.world * self.cap
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(self.lib.mucon_peer_alloc(C.c_size_t(nbytes), C.byref(ptr), handle), "mucon_peer_alloc")
        self._own = ptr.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy().reshape(self.world, 64)
        self._opened, self.ptrs = [], []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self._own)
                continue
            q = C.c_void_p()
            hb = (C.c_ubyte * 64)(*allh[r].tolist())
            _lib.check(self.lib.mucon_peer_open(hb, C.byref(q)), "mucon_peer_open")
            self._opened.append(q.value)
            self.ptrs.append(q.value)
        self.recv = torch.as_tensor(_DevBuf(self._own, nbytes), device=dev).view(len(self.plans), self.world, self.cap)
        for slot, p in enumerate(self.plans):
            p.peer_delta = peer_deltas(p.payload.data_ptr(), self.ptrs, self.rank, self.world, slot, self.cap)
        dist.barrier(group=group)

    def plan(self, i):
        return self.plans[i % len(self.plans)]

    def fence(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def result(self, i):
        return self.recv[i % len(self.plans)]

    def close(self):
        for p in self.plans:
            p.peer_delta = None
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for q in self._opened:
            self.lib.mucon_peer_close(C.c_void_p(q))
        self._opened = []
        dist.barrier(group=self.group)
        if self._own:
            self.recv = None
            self.lib.mucon_peer_free(C.c_void_p(self._own))
            self._own = None
