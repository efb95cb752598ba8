"""Transcript-constrained Viterbi alignment on B200 -- host side.

Drop-in surface (reference src/core/viterbi/viterbi.py:34-65, used by
src/mucon/evaluators.py:80,148,167,178-180):

    dec = Viterbi(grammar, length_model, frame_sampling=30)
    dec.grammar = ...; dec.length_model = ...          # re-assigned per video by the evaluator
    score, labels, segments = dec.decode(log_frame_probs)   # np.ndarray [T, C]

plus a batched engine (``AlignPlan`` / ``ViterbiEngine``) that decodes many (video, candidate
transcript) units in two kernel launches; that is what bench.py and the multi-GPU driver use.
All compute happens in libmucon_b200.so (CUDA, sm_100a); there is no CPU fallback.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from .grammar import lower_grammar, tie_ranks, tie_ranks_for_lists
from .length_model import PoissonModel, log_factorial_prefix

__all__ = ["Viterbi", "ViterbiEngine", "AlignPlan", "Segment", "default_seg0_f32"]


_NUMPY2 = int(np.__version__.split(".")[0]) >= 2


def default_seg0_f32(dtype):
    """The float mix the installed NumPy gives the reference decoder:
    with float32 log-probs NumPy >= 2 keeps segment 0 in float32, NumPy 1.x promotes to float64."""
    return _NUMPY2 and np.dtype(dtype) == np.float32


def _raw_stream(device):
    """cudaStream_t of torch's current stream on `device` as an int (the private fast path triton also uses; the
    public objects cost two extra microseconds per call, which the one-video-per-call pattern notices)."""
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream(device).cuda_stream


class Segment(object):
    """Same attributes as the reference's traceback Segment (viterbi.py:141-143)."""
    __slots__ = ("label", "length")

    def __init__(self, label, length):
        self.label, self.length = label, length

    def __repr__(self):
        return f"Segment(label={self.label}, length={self.length})"


def _align16(n):
    return (n + 15) & ~15


class _Blob:
    """Packs several small host arrays into one buffer so they reach the GPU in one copy."""

    def __init__(self):
        self.parts = []
        self.size = 0

    def add(self, name, arr):
        arr = np.ascontiguousarray(arr)
        off = self.size
        self.parts.append((name, off, arr))
        self.size = _align16(off + arr.nbytes)
        return off

    def upload(self, device, stream=None):
        # pinning costs a cudaHostAlloc (~100 us): worth it for a batch plan, not for one video
        host = torch.empty(max(self.size, 16), dtype=torch.uint8,
                           pin_memory=torch.cuda.is_available() and self.size >= (1 << 18))
        hv = host.numpy()
        for _, off, arr in self.parts:
            hv[off:off + arr.nbytes] = arr.view(np.uint8).reshape(-1)
        dev = host.to(device, non_blocking=True)
        base = dev.data_ptr()
        return dev, host, {name: base + off for name, off, _ in self.parts}


class FlatCandidates:
    """Candidate transcripts of a batch as three arrays: n_cands [V] (candidates per video), lengths [U]
    (labels per candidate, videos in order) and labels [sum lengths] (all transcripts concatenated).
    tie_rank [U, 2] (optional, grammar.tie_ranks): which candidate wins an EXACT score tie the way the reference's
    hypothesis dict decides it; without it the lowest candidate index wins."""

    def __init__(self, n_cands, lengths, labels, tie_rank=None):
        self.n_cands, self.lengths, self.labels, self.tie_rank = n_cands, lengths, labels, tie_rank

    @classmethod
    def from_lists(cls, candidates):
        lengths = np.fromiter((len(t) for cl in candidates for t in cl), dtype=np.int64)
        labels = np.fromiter((x for cl in candidates for t in cl for x in t), dtype=np.int32, count=int(lengths.sum()))
        ties = np.concatenate([tie_ranks_for_lists(cl) for cl in candidates]) if len(candidates) else None
        return cls(np.fromiter((len(cl) for cl in candidates), dtype=np.int64, count=len(candidates)), lengths, labels,
                   ties)


MAX_J_REGISTER = 128  # kDpMaxJ in csrc/viterbi_dp.cuh
MAX_N_REGISTER = 65   # dp_max_n(8)
LONG_TAIL_FRACTION = float(os.environ.get("MUCON_LONG_TAIL_FRACTION", "0.85"))
LONG_TAIL_MAX_UNITS = 40
LANES_MIN_WARPS = 1184  # two warps per SM sub-partition on a 148-SM part


class AlignPlan:
    """Shapes, offsets and device metadata of one batch of (video, candidate) units.

    T            frames per video [V]
    candidates   per video, a list of candidate transcripts (list of int lists)
    len_params   per video [C, 3] float64 (ln m, m, norms) -- Poisson fast path, or
    len_rows     per unit [N_u, J] float64 length scores for j = 1..J blocks (any length model)
    """

    def __init__(self, T, candidates, n_classes, fs=30, max_len=2000, len_params=None, len_rows=None,
                 device=None, want_bp=True, labels="best", groups=None, long_K=None, payload_capacity=None,
                 force_generic=False, len_params_dev=None, tie_rank=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.fs, self.max_len, self.C = int(fs), int(max_len), int(n_classes)
        self.J = self.max_len // self.fs
        T = np.asarray(T, dtype=np.int64)
        self.T = T
        self.V = V = int(T.shape[0])
        self.K = K = T // self.fs
        self.vid_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
        self.blk_off = np.concatenate([[0], np.cumsum(K)]).astype(np.int64)
        if isinstance(candidates, FlatCandidates):
            # arrays in, no per-transcript Python objects (a 100k-unit plan builds in tens of milliseconds)
            ncand = np.asarray(candidates.n_cands, dtype=np.int64)
            nlen = np.asarray(candidates.lengths, dtype=np.int64)
            tr_all = np.ascontiguousarray(candidates.labels, dtype=np.int32).reshape(-1)
            if ncand.shape[0] != V or (V and ncand.min() < 1):
                raise ValueError("need at least one candidate transcript per video")
            if int(ncand.sum()) != nlen.shape[0] or int(nlen.sum()) != tr_all.shape[0]:
                raise ValueError("FlatCandidates: n_cands / lengths / labels do not add up")
        else:
            ncand = np.array([len(c) for c in candidates], dtype=np.int64)
            if len(candidates) != V or (V and ncand.min() < 1):
                raise ValueError("need at least one candidate transcript per video")
            flat = [np.asarray(tr, dtype=np.int32).reshape(-1) for cl in candidates for tr in cl]
            nlen = np.array([t.shape[0] for t in flat], dtype=np.int64)
            tr_all = np.concatenate(flat).astype(np.int32) if flat else np.zeros(0, np.int32)
        self.cand_off = np.concatenate([[0], np.cumsum(ncand)]).astype(np.int32)
        self.U = U = int(self.cand_off[-1])
        self.unit_vid = np.repeat(np.arange(V, dtype=np.int32), ncand)
        if U and nlen.min() < 1:
            raise ValueError("empty transcript")  # the reference crashes at one_hot (SURVEY V-edge)
        self.N = nlen
        self.tr_off = np.concatenate([[0], np.cumsum(nlen)]).astype(np.int32)
        self.tr = tr_all
        if U and (self.tr.min() < 0 or self.tr.max() >= self.C):
            raise ValueError("transcript label outside [0, n_classes)")
        self.max_N = int(nlen.max()) if U else 1
        self.single = bool(U == V)
        self.labels_mode = labels
        # which candidate wins an EXACT score tie: the reference's dict order (mucon_viterbi_select_ranked).  Lists of
        # transcripts get the ranks ModifiedPathGrammar would imply; FlatCandidates carry their own (`tie_rank`
        # attribute) or fall back to "lowest index" (mucon_viterbi_select).
        if tie_rank is None and isinstance(candidates, FlatCandidates):
            tie_rank = getattr(candidates, "tie_rank", None)
        elif tie_rank is None and not self.single and labels == "best":
            tie_rank = np.concatenate([tie_ranks_for_lists(cl) for cl in candidates]) if U else None
        if tie_rank is not None:
            tie_rank = np.ascontiguousarray(tie_rank, dtype=np.int32)
            if tie_rank.shape != (U, 2):
                raise ValueError("tie_rank must be [n_units, 2]")
        self.tie_rank = tie_rank
        uT = T[self.unit_vid]
        uK = K[self.unit_vid]
        # labels: per-unit ("all") or per-video, written for the best candidate ("best")
        if labels == "all":
            self.lab_off = np.concatenate([[0], np.cumsum(uT)]).astype(np.int64)[:-1]
            self.n_labels = int(uT.sum())
        elif self.single:
            self.lab_off = self.vid_off[:-1].copy()
            self.n_labels = int(T.sum())
        else:
            self.lab_off = np.full(U, -1, dtype=np.int64)
            self.n_labels = int(T.sum())
        bp_sz = uK * nlen
        self.bp_off = np.concatenate([[0], np.cumsum(bp_sz)]).astype(np.int64)
        self.n_bp = int(self.bp_off[-1])
        self.total_frames = int(T.sum())
        self.total_blocks = int(K.sum())
        self.aligned_frames = int(uT.sum())  # the benchmark's unit of work: T x candidates
        # Launch plan.  Videos are sorted longest first and cut into groups; group g's DP runs on
        # a second stream while group g+1 is still being scanned (the scan is HBM-bound, the DP is
        # issue-bound, so they overlap well).  Group 0 holds the longest videos: their DP is a long
        # serial chain, so it starts first and gives every transcript segment its own warp.
        self.order_v = np.argsort(-T, kind="stable").astype(np.int32)
        self.max_K = int(uK.max()) if U else 0
        # shapes outside the register-resident kernels (J > 128, e.g. the reference class's default
        # frame_sampling = 1, or N > 65) run the generic kernel with its state in a workspace
        self.generic = bool(force_generic) or self.J > MAX_J_REGISTER or self.max_N > MAX_N_REGISTER
        if groups is None:
            groups = 1  # measured: cutting the scan into groups serialises its long videos
        cumT = np.cumsum(T[self.order_v]) / max(1, T.sum())
        cuts = [0.12, 0.40, 0.72, 1.0] if groups == 4 else list(np.linspace(0, 1, groups + 1)[1:])
        bounds = [0] + [int(np.searchsorted(cumT, c, side="left")) + 1 for c in cuts[:-1]] + [V]
        bounds = sorted(set(min(max(x, 0), V) for x in bounds))
        lib = _lib.lib()
        n32 = np.ascontiguousarray(nlen, dtype=np.int32)
        group_of_video = np.zeros(V, dtype=np.int64)
        for gi in range(len(bounds) - 1):
            group_of_video[self.order_v[bounds[gi]:bounds[gi + 1]]] = gi
        unit_group = group_of_video[self.unit_vid]
        # units longest first (by blocks, then segments): the order of the packers and of the fused launch
        order_all = np.argsort(-(uK * 1024 + nlen), kind="stable").astype(np.int32)
        self.groups = []
        wu_all = []
        for gi in range(len(bounds) - 1):
            v0, v1 = bounds[gi], bounds[gi + 1]
            units = np.nonzero(unit_group == gi)[0]
            if v1 <= v0:
                continue
            order_u = order_all if len(bounds) == 2 else \
                units[np.argsort(-(uK[units] * 1024 + nlen[units]), kind="stable")].astype(np.int32)
            gmaxN = int(nlen[units].max()) if units.size else 1
            gmaxK = int(uK[units].max()) if units.size else 0
            want = 32 if (gi == 0 and len(bounds) > 2 and long_K and gmaxK >= long_K) else 0
            wu = np.full(max(units.size, 1) * 16, -1, dtype=np.int32)
            n_cta, wpc, lanes = C.c_int32(0), C.c_int32(4), C.c_int32(0)
            if not self.generic:
                _lib.check(lib.mucon_viterbi_pack_h(
                    n32.ctypes.data_as(C.c_void_p), order_u.ctypes.data_as(C.c_void_p), C.c_int(int(units.size)),
                    C.c_int(gmaxN), C.c_int(self.fs), C.c_int(self.max_len), C.c_int(want),
                    wu.ctypes.data_as(C.c_void_p), C.byref(n_cta), C.byref(wpc), C.byref(lanes)),
                    "mucon_viterbi_pack_h")
            self.groups.append(dict(v0=v0, v1=v1, n_cta=int(n_cta.value), wpc=int(wpc.value), lanes=int(lanes.value),
                                    max_N=gmaxN, max_K=gmaxK, wu_off=sum(len(x) for x in wu_all)))
            wu_all.append(wu[:int(n_cta.value) * int(wpc.value)])
        self.warp_unit = np.concatenate(wu_all) if wu_all else np.full(16, -1, np.int32)
        if self.warp_unit.size == 0:
            self.warp_unit = np.full(16, -1, np.int32)
        self.n_cta = sum(g["n_cta"] for g in self.groups)
        self.wpc = self.groups[0]["wpc"] if self.groups else 4

        # lane-per-segment packing (J <= 66, N <= 33): see csrc/viterbi_lanes.cuh
        self.n_lane_warps = 0
        self.lane_unit = None
        if U and not self.generic and self.J <= 66 and self.max_N <= 33:
            lu = np.full(U * 32, -1, dtype=np.int32)
            nw = C.c_int32(0)
            _lib.check(lib.mucon_viterbi_pack_lanes_h(
                n32.ctypes.data_as(C.c_void_p), order_all.ctypes.data_as(C.c_void_p), C.c_int(U),
                lu.ctypes.data_as(C.c_void_p), C.byref(nw)), "mucon_viterbi_pack_lanes_h")
            self.n_lane_warps = int(nw.value)
            self.lane_unit = lu[:self.n_lane_warps * 32]

        blob = _Blob()
        blob.add("vid_off", self.vid_off)
        blob.add("blk_off", self.blk_off)
        blob.add("unit_vid", self.unit_vid)
        blob.add("tr", self.tr)
        blob.add("tr_off", self.tr_off)
        blob.add("cand_off", self.cand_off)
        if self.tie_rank is not None:
            blob.add("tie_rank", self.tie_rank.reshape(-1))
        blob.add("lab_off", self.lab_off)
        blob.add("bp_off", self.bp_off[:-1] if U else self.bp_off)
        blob.add("order_v", self.order_v)
        blob.add("warp_unit", self.warp_unit)
        blob.add("order_u", order_all)
        # long_K: videos of >= long_K blocks go to a wide launch of their own, a warp per transcript segment
        # (None = choose automatically, 0 = never)
        if long_K is None:
            # auto: the few videos within ~15 % of the longest one set the critical path of the launch
            # (their DP is a serial chain of K steps); they get a wide launch of their own, a warp per
            # transcript segment, concurrent with the main one (measured on the 1712-video split:
            # 171 -> 155 us with the 22 longest videos split off; more than ~40 and it stops paying)
            long_K = 0
            if U >= 256 and self.single and self.max_N <= 15 and self.max_K >= 128:
                thr = max(1, int(LONG_TAIL_FRACTION * self.max_K))
                if int((uK >= thr).sum()) <= LONG_TAIL_MAX_UNITS:
                    long_K = thr
        self.n_long = int((uK >= long_K).sum()) if long_K else 0
        blob.add("vid_lab_off", self.vid_off[:-1])
        if self.lane_unit is not None:
            blob.add("lane_unit", self.lane_unit)
        if self.generic:
            blob.add("ws_off", (2 * self.J * self.tr_off[:-1].astype(np.int64)) if U else np.zeros(1, np.int64))
        self.use_rows = len_rows is not None
        if self.use_rows:
            rows = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, self.J) for r in len_rows]) \
                if U else np.zeros((0, self.J))
            if rows.shape[0] != self.tr.shape[0]:
                raise ValueError("len_rows must hold one [N_u, J] block per unit")
            blob.add("len_rows", rows)
        else:
            if len_params is None and len_params_dev is None:
                raise ValueError("need len_params (Poisson) or len_rows")
            if len_params_dev is None:
                lp = np.asarray(len_params, dtype=np.float64).reshape(V, self.C, 3)
                pos_vid = np.repeat(self.unit_vid, nlen)
                blob.add("len_params", lp[pos_vid, self.tr])
            lf = log_factorial_prefix(self.max_len - 1)
            idx = np.minimum(np.arange(self.J + 1) * self.fs, self.max_len - 1)
            blob.add("logfact", lf[idx])
        self.h2d_meta_bytes = blob.size
        self._meta_dev, self._meta_host, self.p = blob.upload(self.device)
        if len_params_dev is not None and not self.use_rows:
            # per-position parameters already on the device (evaluate.class_mean_params_device): [sum N, 3] float64
            if (not len_params_dev.is_cuda or len_params_dev.dtype != torch.float64 or not len_params_dev.is_contiguous()
                    or len_params_dev.numel() != 3 * int(self.tr_off[-1])):
                raise ValueError("len_params_dev must be a contiguous float64 CUDA tensor [sum N, 3]")
            self._len_params_dev = len_params_dev
            self.p["len_params"] = len_params_dev.data_ptr()

        dev = self.device
        # scores and segment lengths -- what the multi-GPU gather moves -- share one buffer so that
        # the collective needs no packing step: [score f64 x U | seg_blocks i32 x sum N | padding]
        n_pos = int(self.tr_off[-1])
        need = 8 * U + 4 * n_pos
        cap = max(need, int(payload_capacity or 0), 16)
        self.payload = torch.zeros((cap + 7) // 8 * 8, dtype=torch.uint8, device=dev)
        self.score = self.payload[:8 * U].view(torch.float64)
        self.final_j = torch.empty(U, dtype=torch.int32, device=dev)
        self.status = torch.empty(U, dtype=torch.int32, device=dev)
        self.seg_blocks = self.payload[8 * U:8 * U + 4 * n_pos].view(torch.int32)
        self.labels = torch.empty(self.n_labels, dtype=torch.int32, device=dev)
        self.bp_u16 = self.J > 255
        self.bp = torch.empty(max(self.n_bp, 1), dtype=torch.uint16 if self.bp_u16 else torch.uint8, device=dev)
        self.ws = torch.empty(max(1, 2 * n_pos * self.J), dtype=torch.float64, device=dev) if self.generic else None
        self.best = torch.empty(V, dtype=torch.int32, device=dev)
        self.bs = None  # allocated by the engine once the input dtype is known


class ViterbiEngine:
    """Runs AlignPlans: block-score scan + DP/traceback/labels (+ candidate arg-max)."""

    def __init__(self, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = _lib.lib()
        self.launches = 0  # kernels launched so far (bench.py reports this)
        self._side = None
        self._events = {}

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _event(self, i):
        ev = self._events.get(i)
        if ev is None:
            ev = self._events[i] = torch.cuda.Event()
        return ev

    def run(self, plan, logp, seg0_f32=None, stream=None, mid_event=None, mode="auto", write_bs=True, z_off=None):
        """logp: CUDA tensor [sum T, C] float32 or float64, videos concatenated.  Asynchronous.

        z_off: int64 CUDA tensor [V+1] -- `logp` is then the table at the backbone's POOLED resolution
        ([sum Tz, C]) and frame t of a video reads row min(floor(t * (float)Tz / T), Tz - 1) (nearest
        interpolate, models.py:574-576); bit-identical to running on the expanded array, fused mode only.

        mode: "fused"  one launch, scan and DP of a video in the same CTA (one transcript per video)
              "split"  block-score scan kernel + DP kernel (any number of candidates per video)
              "auto"   fused when every video has a single candidate and the shape fits
        write_bs: fused mode only -- also store the block scores to plan.bs."""
        if not logp.is_cuda:
            raise _lib.MuconError("ViterbiEngine.run needs a CUDA tensor (no CPU fallback)")
        if logp.dtype not in (torch.float32, torch.float64) or not logp.is_contiguous():
            raise TypeError("logp must be contiguous float32/float64")
        if z_off is not None:
            if mode not in ("auto", "fused") or not plan.single or plan.generic:
                raise _lib.MuconError("a pooled-resolution source needs the fused kernel (one transcript per video)")
            if z_off.dtype != torch.int64 or z_off.numel() != plan.V + 1 or not z_off.is_cuda or logp.shape[1] != plan.C:
                raise ValueError("z_off must be an int64 CUDA tensor [V+1]; logp [sum Tz, C]")
            mode = "fused"
        elif logp.shape[0] != plan.total_frames or logp.shape[1] != plan.C:
            raise ValueError(f"logp shape {tuple(logp.shape)} != ({plan.total_frames}, {plan.C})")
        is64 = logp.dtype == torch.float64
        if seg0_f32 is None:
            seg0_f32 = default_seg0_f32(np.float64 if is64 else np.float32)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        sp = C.c_void_p(st.cuda_stream)
        if plan.bs is None or plan.bs.dtype != logp.dtype:
            plan.bs = torch.empty((max(plan.total_blocks, 1), plan.C), dtype=logp.dtype, device=self.device)
        p = plan.p
        if plan.V == 0:
            return plan
        lib = self.lib
        b = _lib.ViterbiBatch()
        b.U, b.C, b.fs, b.max_len = plan.U, plan.C, plan.fs, plan.max_len
        b.bs_is_f64, b.seg0_f32 = int(is64), int(bool(seg0_f32))
        b.bs = plan.bs.data_ptr()
        b.vid_off, b.blk_off, b.unit_vid = p["vid_off"], p["blk_off"], p["unit_vid"]
        b.tr, b.tr_off = p["tr"], p["tr_off"]
        if plan.use_rows:
            b.len_rows, b.len_params, b.logfact = p["len_rows"], None, None
        else:
            b.len_rows, b.len_params, b.logfact = None, p["len_params"], p["logfact"]
        b.lab_off, b.bp_off = p["lab_off"], p["bp_off"]
        b.score, b.labels = plan.score.data_ptr(), plan.labels.data_ptr()
        b.seg_blocks, b.bp = plan.seg_blocks.data_ptr(), plan.bp.data_ptr()
        b.final_j, b.status = plan.final_j.data_ptr(), plan.status.data_ptr()
        deltas = getattr(plan, "peer_delta", None)   # dist.PeerExchange: result stores repeated into peers' buffers
        if deltas:
            b.n_peers = len(deltas)
            for i_, d_ in enumerate(deltas):
                b.peer_delta[i_] = d_
        self.last_mode = "split"
        if mode not in ("auto", "fused", "split", "lanes"):
            raise ValueError(mode)
        if plan.generic:
            if mode == "fused":
                raise _lib.MuconError("fused mode does not cover J > %d / N > %d" % (MAX_J_REGISTER, MAX_N_REGISTER))
            _lib.check(lib.mucon_viterbi_blockscores(
                _lib.ptr(logp), C.c_int(int(is64)), C.c_void_p(p["vid_off"]), C.c_void_p(p["blk_off"]),
                C.c_void_p(p["order_v"]), C.c_int(plan.V), C.c_int(plan.C), C.c_int(plan.fs),
                _lib.ptr(plan.bs), sp), "mucon_viterbi_blockscores")
            if mid_event is not None:
                mid_event.record(st)
            b.max_N, b.max_K, b.n_cta, b.wpc, b.lanes, b.warp_unit = plan.max_N, plan.max_K, 0, 4, 0, None
            _lib.check(lib.mucon_viterbi_decode_generic(
                C.byref(b), _lib.ptr(plan.ws), C.c_void_p(p["ws_off"]), C.c_int(int(plan.bp_u16)), sp),
                "mucon_viterbi_decode_generic")
            self.launches += 2
            self.last_mode = "generic"
            return self._finish(plan, sp)
        # candidate sets large enough to keep every scheduler busy: the lane-per-segment kernel
        # needs a third of the instructions (measured 1.66x on 16384 units); small batches are
        # latency-bound and stay on the shift-register kernel
        if mode == "auto" and not plan.single and plan.lane_unit is not None and plan.n_lane_warps >= LANES_MIN_WARPS:
            mode = "lanes"
        if mode == "lanes":
            if plan.lane_unit is None:
                raise _lib.MuconError("lanes mode needs J <= 66 and N <= 33")
            _lib.check(lib.mucon_viterbi_blockscores(
                _lib.ptr(logp), C.c_int(int(is64)), C.c_void_p(p["vid_off"]), C.c_void_p(p["blk_off"]),
                C.c_void_p(p["order_v"]), C.c_int(plan.V), C.c_int(plan.C), C.c_int(plan.fs),
                _lib.ptr(plan.bs), sp), "mucon_viterbi_blockscores")
            if mid_event is not None:
                mid_event.record(st)
            b.max_N, b.max_K, b.n_cta, b.wpc, b.lanes, b.warp_unit = plan.max_N, plan.max_K, 0, 4, 0, None
            _lib.check(lib.mucon_viterbi_decode_lanes(
                C.byref(b), C.c_void_p(p["lane_unit"]), C.c_int(plan.n_lane_warps), None, sp),
                "mucon_viterbi_decode_lanes")
            self.launches += 2
            self.last_mode = "lanes"
            return self._finish(plan, sp)
        if mode == "fused" or (mode == "auto" and plan.single):
            # One call: the long tail (plan.n_long longest videos, first in order_u) goes to a wide
            # launch, everything else to the main launch, which follows in the same stream as a
            # programmatic dependent launch and runs concurrently with it.
            b.n_cta, b.wpc, b.warp_unit = 0, 4, None
            b.U, b.lanes, b.max_N, b.max_K = plan.U, 0, plan.max_N, plan.max_K
            n_long = plan.n_long if (plan.max_N <= 15 and plan.U >= 64) else 0
            if z_off is not None:
                rc = lib.mucon_viterbi_align_fused_pooled(
                    C.byref(b), _lib.ptr(logp), C.c_int(int(is64)), _lib.ptr(z_off), C.c_void_p(p["order_u"]),
                    C.c_int(n_long), C.c_int(int(bool(write_bs))), sp)
            else:
                rc = lib.mucon_viterbi_align_fused_tail(
                    C.byref(b), _lib.ptr(logp), C.c_int(int(is64)), C.c_void_p(p["order_u"]), C.c_int(n_long),
                    C.c_int(int(bool(write_bs))), sp)
            if rc == 0:
                self.launches += 2 if n_long else 1
                self.last_mode = "fused"
                if mid_event is not None:
                    mid_event.record(st)
                return self._finish(plan, sp)
            if rc != -2 or mode == "fused":
                _lib.check(rc, "mucon_viterbi_align_fused")
            if z_off is not None:
                raise _lib.MuconError("shape not covered by the fused kernel: expand the log-probabilities instead")
        overlap = len(plan.groups) > 1
        if overlap:
            side = self._side_stream()
            ssp = C.c_void_p(side.cuda_stream)
        for gi, g in enumerate(plan.groups):
            _lib.check(lib.mucon_viterbi_blockscores(
                _lib.ptr(logp), C.c_int(int(is64)), C.c_void_p(p["vid_off"]), C.c_void_p(p["blk_off"]),
                C.c_void_p(p["order_v"] + 4 * g["v0"]), C.c_int(g["v1"] - g["v0"]), C.c_int(plan.C), C.c_int(plan.fs),
                _lib.ptr(plan.bs), sp), "mucon_viterbi_blockscores")
            if mid_event is not None and gi == len(plan.groups) - 1:
                mid_event.record(st)  # end of the last scan: lets bench.py split scan and DP tail
            b.max_N, b.max_K, b.n_cta, b.wpc, b.lanes = g["max_N"], g["max_K"], g["n_cta"], g["wpc"], g["lanes"]
            b.warp_unit = p["warp_unit"] + 4 * g["wu_off"]
            if overlap:
                ev = self._event(gi)
                ev.record(st)
                side.wait_event(ev)
                _lib.check(lib.mucon_viterbi_decode(C.byref(b), ssp), "mucon_viterbi_decode")
            else:
                _lib.check(lib.mucon_viterbi_decode(C.byref(b), sp), "mucon_viterbi_decode")
            self.launches += 2
        if overlap:
            done = self._event(len(plan.groups))
            done.record(side)
            st.wait_event(done)
        return self._finish(plan, sp)

    def _finish(self, plan, sp):
        lib, p = self.lib, plan.p
        if plan.labels_mode == "best" and not plan.single:
            if plan.tie_rank is not None:
                _lib.check(lib.mucon_viterbi_select_ranked(
                    _lib.ptr(plan.score), _lib.ptr(plan.status), C.c_void_p(p["cand_off"]), C.c_int(plan.V),
                    _lib.ptr(plan.final_j), C.c_void_p(p["tr_off"]), C.c_void_p(p["tie_rank"]),
                    _lib.ptr(plan.best), sp), "mucon_viterbi_select_ranked")
            else:
                _lib.check(lib.mucon_viterbi_select(
                    _lib.ptr(plan.score), _lib.ptr(plan.status), C.c_void_p(p["cand_off"]), C.c_int(plan.V),
                    _lib.ptr(plan.best), sp), "mucon_viterbi_select")
            _lib.check(lib.mucon_viterbi_labels(
                _lib.ptr(plan.best), C.c_int(plan.V), C.c_void_p(p["vid_lab_off"]), C.c_void_p(p["vid_off"]),
                C.c_void_p(p["unit_vid"]), C.c_void_p(p["tr"]), C.c_void_p(p["tr_off"]),
                _lib.ptr(plan.seg_blocks), C.c_int(plan.fs), _lib.ptr(plan.labels), sp), "mucon_viterbi_labels")
            self.launches += 2
        return plan

    # ---- host-side conveniences ------------------------------------------------------------
    def fetch(self, plan, want_bp=False):
        """Synchronising device->host read of a finished plan's results (numpy)."""
        out = dict(score=plan.score.cpu().numpy(), final_j=plan.final_j.cpu().numpy(),
                   status=plan.status.cpu().numpy(), seg_blocks=plan.seg_blocks.cpu().numpy(),
                   labels=plan.labels.cpu().numpy())
        if not plan.single and plan.labels_mode == "best":
            out["best"] = plan.best.cpu().numpy()
        if want_bp:
            out["bp"] = plan.bp.cpu().numpy()[:plan.n_bp]
            out["bs"] = plan.bs.cpu().numpy()[:plan.total_blocks]
        return out


class HostAlignPipeline:
    """Alignment of a STREAM of host batches (pinned log-probabilities in, pinned labels / scores / segment lengths out)
    with the three legs of a step on three streams: the host->device copy of batch i + 1 runs while the kernels of
    batch i run and while the results of batch i go back (PCIe is full duplex), so a steady stream costs the
    host->device copy alone.  `depth` device input buffers; a slot is reused only after its results were collected.

        pipe = HostAlignPipeline(engine)
        t = pipe.submit(host_logp, lambda: AlignPlan(T, candidates, C, device=engine.device, len_params=...))
        plan, labels, score, seg_blocks = pipe.result(t)      # pinned host tensors, valid until the slot is reused
    """

    def __init__(self, engine, depth=2):
        self.eng, self.depth = engine, int(depth)
        dev = engine.device
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots = [dict(busy=False) for _ in range(self.depth)]
        self.n = 0
        self.h2d_bytes = self.d2h_bytes = 0

    @staticmethod
    def _fit(slot, name, numel, dtype, **kw):
        t = slot.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = slot[name] = torch.empty(max(int(numel), 1), dtype=dtype, **kw)
        return t[:numel]

    def submit(self, host_logp, make_plan, seg0_f32=None):
        """host_logp: pinned CPU tensor [sum T, C] (float32 / float64); make_plan(): the batch's AlignPlan -- called here,
        on the host, while the copy is in flight.  Returns a ticket for result()."""
        if host_logp.is_cuda or not host_logp.is_contiguous():
            raise TypeError("host_logp must be a contiguous CPU tensor (pinned for an asynchronous copy)")
        ticket = self.n
        slot = self.slots[ticket % self.depth]
        if slot["busy"]:
            raise _lib.MuconError("HostAlignPipeline: collect result(%d) before submitting again" % slot["ticket"])
        dev = self.eng.device
        dev_in = slot.get("dev_in")
        if dev_in is None or dev_in.shape != host_logp.shape or dev_in.dtype != host_logp.dtype:
            dev_in = slot["dev_in"] = torch.empty(host_logp.shape, dtype=host_logp.dtype, device=dev)
        cur = torch.cuda.current_stream(dev)
        self.s_in.wait_stream(cur)            # whatever produced host_logp / freed the buffers on the caller's stream
        if slot.get("ran") is not None:
            self.s_in.wait_event(slot["ran"])  # the previous kernels of this slot have read dev_in
        with torch.cuda.stream(self.s_in):
            dev_in.copy_(host_logp, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        with torch.cuda.stream(self.s_run):
            plan = make_plan()                 # host work + the plan's metadata upload, under the copy
            self.s_run.wait_event(ev_in)
            self.eng.run(plan, dev_in, seg0_f32=seg0_f32, write_bs=False, stream=self.s_run)
            ran = torch.cuda.Event()
            ran.record(self.s_run)
        self.s_out.wait_event(ran)
        with torch.cuda.stream(self.s_out):
            labels = self._fit(slot, "labels", plan.labels.numel(), torch.int32, pin_memory=True)
            score = self._fit(slot, "score", plan.score.numel(), torch.float64, pin_memory=True)
            seg = self._fit(slot, "seg", plan.seg_blocks.numel(), torch.int32, pin_memory=True)
            labels.copy_(plan.labels, non_blocking=True)
            score.copy_(plan.score, non_blocking=True)
            seg.copy_(plan.seg_blocks, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.s_out)
        slot.update(busy=True, ticket=ticket, plan=plan, ran=ran, done=done, out=(labels, score, seg))
        self.h2d_bytes = host_logp.numel() * host_logp.element_size() + getattr(plan, "h2d_meta_bytes", 0)
        self.d2h_bytes = labels.numel() * 4 + score.numel() * 8 + seg.numel() * 4
        self.n += 1
        return ticket

    def result(self, ticket):
        """Blocks until batch `ticket` is back on the host: (plan, labels, score, seg_blocks)."""
        slot = self.slots[ticket % self.depth]
        if not slot["busy"] or slot["ticket"] != ticket:
            raise _lib.MuconError("HostAlignPipeline: no pending batch %d" % ticket)
        slot["done"].synchronize()
        slot["busy"] = False
        return (slot["plan"],) + slot["out"]


def _length_rows(length_model, transcript, fs, J):
    """Generic lowering of any LengthModel: rows[n, j-1] = score(j*fs, label_n)."""
    tab = getattr(length_model, "poisson", None)
    max_len = length_model.max_length()
    rows = np.full((len(transcript), J), -np.inf, dtype=np.float64)
    for j in range(1, J + 1):
        l = j * fs
        if tab is not None and isinstance(tab, np.ndarray):
            if l < max_len:
                rows[:, j - 1] = tab[l, transcript]
        else:
            rows[:, j - 1] = [length_model.score(l, int(c)) for c in transcript]
    return rows


class Viterbi(object):
    """Drop-in for core.viterbi.viterbi.Viterbi (viterbi.py:10-65), decoding on the GPU.

    grammar: SingleTranscriptGrammar / PathGrammar / ModifiedPathGrammar (this package's or the
    reference's).  length_model: PoissonModel (this package's: parameter fast path; the
    reference's: rows taken from its table) or any LengthModel (rows from score()).
    ``np_mode``: None = behave like the installed NumPy; "numpy1" / "numpy2" force the float
    promotion regime of SURVEY.md section 0.4.
    """

    def __init__(self, grammar, length_model, frame_sampling=1, max_hypotheses=np.inf, device=None, np_mode=None):
        if max_hypotheses != np.inf:
            raise NotImplementedError("hypothesis pruning is not implemented (the evaluator never prunes)")
        self.grammar = grammar
        self.length_model = length_model
        self.frame_sampling = frame_sampling
        self.max_hypotheses = max_hypotheses
        self.np_mode = np_mode
        # one transcript + PoissonModel: the whole decode is one C call (mucon_single_decode_h).  Set to False to
        # go through the general AlignPlan path, which keeps the raw tables (back-pointers, block scores) for `last`.
        self.fast_single = True
        self._engine = None
        self._device = device
        self._last = None

    @property
    def last(self):
        """Raw outputs of the last decode (scores, segment lengths, status, labels, back-pointer table,
        block scores), fetched from the device on demand."""
        if self._last is None:
            return None
        eng, plan, out = self._last
        if "bp" not in out:
            out.update(eng.fetch(plan, want_bp=True))
        return out

    def set_multi_length(self, mode=True):  # viterbi.py:40-41 -- a no-op there too
        pass

    def _eng(self):
        if self._engine is None:
            self._engine = ViterbiEngine(self._device)
        return self._engine

    def _decode_single(self, eng, logp, tr, lm, fs, max_len):
        """One transcript, Poisson lengths: the whole decode is one C call (mucon_single_decode_h: staging copy,
        fused kernel, result copy).  Returns None when the shape needs the general path."""
        T, Cn = logp.shape
        N = len(tr)
        is64 = logp.dtype == np.float64
        lib = eng.lib
        ses = getattr(eng, "_single", None)
        if ses is None or ses[1] < T or ses[2] != Cn or ses[3] < N or ses[4] != is64:
            if ses is not None:
                lib.mucon_single_destroy(ses[0])
            h = C.c_void_p()
            cap_T, cap_N = max(T, 4096, 2 * ses[1] if ses and ses[1] < T else 0), max(N, 16)
            if eng.device.index is not None:
                torch.cuda.set_device(eng.device)
            _lib.check(lib.mucon_single_create(C.c_int(cap_T), C.c_int(Cn), C.c_int(cap_N), C.c_int(8 if is64 else 4),
                                               C.byref(h)), "mucon_single_create")
            bufs = (np.empty(cap_T, np.int32), np.empty(cap_N, np.int32), np.zeros(2, np.float64), np.zeros(2, np.int32))
            # the result buffers live as long as the session: their addresses are taken once (an ndarray.ctypes access
            # builds a helper object, about a microsecond each -- eight of them per decode otherwise)
            ptrs = (C.c_void_p(bufs[2].ctypes.data), C.c_void_p(bufs[0].ctypes.data), C.c_void_p(bufs[1].ctypes.data),
                    C.c_void_p(bufs[3].ctypes.data), C.c_void_p(bufs[3].ctypes.data + 4))
            ses = eng._single = (h, cap_T, Cn, cap_N, is64) + bufs + (ptrs,)
        h, _, _, _, _, labels, segb, score, st_fj, ptrs = ses
        lp = logp if logp.flags.c_contiguous else np.ascontiguousarray(logp)
        tr32 = np.asarray(tr, dtype=np.int32)
        params = lm.params[tr32]                      # a fancy-indexed copy: C-contiguous [N, 3] float64
        if self.np_mode is None:
            seg0 = default_seg0_f32(logp.dtype)
        else:
            seg0 = logp.dtype == np.float32 and self.np_mode == "numpy2"
        rc = lib.mucon_single_decode_h(
            h, C.c_void_p(lp.ctypes.data), C.c_int(int(is64)), C.c_int(T), C.c_void_p(tr32.ctypes.data), C.c_int(N),
            C.c_void_p(params.ctypes.data), C.c_int(fs), C.c_int(max_len), C.c_int(int(bool(seg0))),
            ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4],
            C.c_void_p(_raw_stream(eng.device)))
        if rc == -2:
            return None
        _lib.check(rc, "mucon_single_decode_h")
        eng.launches += 1
        eng.last_mode = "fused"
        self._last = None
        if st_fj[0] == _lib.UNIT_INFEASIBLE:
            raise AttributeError("no hypothesis survives: sequence too long for this transcript and max_length")
        K = T // fs
        segs = [Segment(l, fs * b) for l, b in zip(tr32.tolist(), segb[:N].tolist()) if b > 0]   # plain ints, no NumPy scalars
        segs[-1].length += T - fs * K
        # list(bytes) hands out CPython's preallocated small ints: about 40 % faster than ndarray.tolist() for 2000 labels
        lab = list(labels[:T].astype(np.uint8).tobytes()) if Cn <= 256 else labels[:T].tolist()
        return np.float64(score[0]), lab, segs

    def decode(self, log_frame_probs):
        logp = np.asarray(log_frame_probs)
        assert logp.shape[1] == self.grammar.n_classes()  # viterbi.py:50
        if logp.dtype not in (np.float32, np.float64):
            logp = logp.astype(np.float64)
        fs = int(self.frame_sampling)
        T, Cn = logp.shape
        if T < fs:
            raise IndexError(f"sequence of {T} frames is shorter than frame_sampling={fs}")
        cands = lower_grammar(self.grammar)
        lm = self.length_model
        max_len = lm.max_length()
        if not math.isfinite(max_len):
            raise TypeError("length model without a finite max_length() is not supported")
        max_len = int(max_len)
        J = max_len // fs
        eng = self._eng()
        if self.fast_single and len(cands) == 1 and isinstance(lm, PoissonModel) and lm.exact_params and J <= MAX_J_REGISTER \
                and Cn % 4 == 0 and Cn <= 128:
            fast = self._decode_single(eng, logp, cands[0], lm, fs, max_len)
            if fast is not None:
                return fast
        kw = {}
        if isinstance(lm, PoissonModel) and lm.exact_params:
            kw["len_params"] = lm.params[None]
        elif isinstance(lm, PoissonModel):
            kw["len_rows"] = [lm.rows_for(tr, fs, J) for tr in cands]   # non-float64 means: the reference's dtypes
        else:
            kw["len_rows"] = [_length_rows(lm, tr, fs, J) for tr in cands]
        if len(cands) > 1 and isinstance(getattr(self.grammar, "successors", None), dict):
            kw["tie_rank"] = tie_ranks(self.grammar.successors, cands, self.grammar.start_symbol())
        plan = AlignPlan([T], [cands], Cn, fs=fs, max_len=max_len, device=eng.device, labels="best", **kw)
        dev_logp = torch.from_numpy(np.ascontiguousarray(logp)).to(eng.device)
        if self.np_mode is None:
            seg0 = default_seg0_f32(logp.dtype)
        else:
            seg0 = logp.dtype == np.float32 and self.np_mode == "numpy2"
        eng.run(plan, dev_logp, seg0_f32=seg0)
        # three device->host reads: [scores | segment lengths] (one buffer), status, labels
        payload = plan.payload.cpu()
        n_pos = int(plan.tr_off[-1])
        out = dict(score=payload[:8 * plan.U].view(torch.float64).numpy(),
                   seg_blocks=payload[8 * plan.U:8 * plan.U + 4 * n_pos].view(torch.int32).numpy(),
                   status=plan.status.cpu().numpy(), labels=plan.labels.cpu().numpy())
        if not plan.single:
            out["best"] = plan.best.cpu().numpy()
        self._last = (eng, plan, out)
        u = 0 if plan.single else int(out["best"][0])
        if u < 0 or out["status"][u] == _lib.UNIT_INFEASIBLE:
            # the reference dies with AttributeError in traceback when every hypothesis has been
            # dropped (K > N*J); same exception type here
            raise AttributeError("no hypothesis survives: sequence too long for this transcript and max_length")
        tr = cands[u]
        sb = out["seg_blocks"][plan.tr_off[u]:plan.tr_off[u + 1]]
        K = T // fs
        segs = [Segment(int(tr[n]), int(fs * sb[n])) for n in range(len(tr)) if sb[n] > 0]
        segs[-1].length += T - fs * K
        return np.float64(out["score"][u]), out["labels"][:T].tolist(), segs
