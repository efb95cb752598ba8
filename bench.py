#!/usr/bin/env python
"""bench.py -- aligned frames/sec of the MuCon Viterbi alignment path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...           (one rank per GPU, NCCL)

Workload (BASELINE.json configs[1], "c2"): a synthetic Breakfast-shaped split of 1712 videos per
GPU (T = clip(lognormal(ln 1800, 0.7), 300, 10000), 48 classes, 2..12 transcript segments,
frame_sampling 30, Poisson length model), float32 log-probabilities.  One step = one pass of the
alignment path over the whole split: the fused scan + DP + traceback + label kernel
(mucon_viterbi_align_fused, one launch), inputs resident in HBM (`value`), or through
the host API with pinned host buffers and H2D/D2H copies inside the timed region (`e2e`).
Weak scaling: every rank aligns its own 1712-video split; for N > 1 the step ends with the NCCL
all_gather of per-video scores and segment lengths.

--impl reference times the reference's CPU algorithm on the host cores: the reference is pure
Python and does not travel to the GPU box, so this is the oracle's hypothesis-table port of it
(oracle/hyp_viterbi.py + the loop-built Poisson table), one process per core.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402  (synthetic input recipes only; no oracle, no CUDA)

C, FS, MAX_LEN, V_PER_GPU = 48, 30, 2000, 1712
METRIC, UNIT = "aligned_frames_per_sec", "frames/s"
WORKLOAD = ("c2: Breakfast-shaped synthetic split, 1712 videos/GPU, T<=10000, 48 classes, N<=12, fs=30, "
            "1 transcript/video, float32 log-probs -> Viterbi alignment (scan + DP + traceback + labels)")


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_split(seed):
    T, trs = synth.breakfast_split(seed=seed, V=V_PER_GPU, C=C, fs=FS, J=MAX_LEN // FS)
    rng = np.random.default_rng(seed + 1000)
    means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, C, int(t))
                      for tr, t in zip(trs, T)])
    return T, trs, means


def device_logp(T, trs, seed, device):
    """log_softmax(N(0,1) + 3*onehot(planted segmentation)), generated on the device."""
    import torch
    rng = np.random.default_rng(seed + 2000)
    planted = np.empty(int(T.sum()), dtype=np.int64)
    pos = 0
    for t, tr in zip(T, trs):
        n = len(tr)
        cuts = np.sort(rng.choice(np.arange(1, t), size=n - 1, replace=False))
        planted[pos:pos + t] = np.repeat(tr, np.diff(np.concatenate([[0], cuts, [t]])))
        pos += t
    g = torch.Generator(device).manual_seed(seed)
    x = torch.randn(int(T.sum()), C, device=device, generator=g)
    x.scatter_add_(1, torch.from_numpy(planted).to(device)[:, None], torch.full((x.shape[0], 1), 3.0, device=device))
    return torch.log_softmax(x, dim=1).contiguous()


def bench_config(world, collective):
    """The `config` object of the JSON line -- the same for both arms (the driver compares them)."""
    T, _, _ = make_split(0)
    Tsum = int(T.sum())
    return {"workload": WORKLOAD, "videos_per_gpu": V_PER_GPU, "frames_per_gpu": Tsum,
            "l2": "inputs (%.0f MB log-probs per GPU) are larger than the 126 MB L2" % (4 * Tsum * C / 1e6),
            "collective": ("none in the step: scores + segment lengths are stored into every rank's receive "
                           "buffer by the kernels (NVLink peer stores); one barrier ends the timed region"
                           if collective == "peer" else
                           "all_gather(scores, segment lengths), overlapped with the next step") if world > 1 else "none",
            "inputs": "host-generated (numpy, seed rank+3000): the same arrays the reference arm samples from; "
                      "every rank has the seed-0 split's lengths / transcripts and its own log-probabilities"}


def make_host_logp(T, trs, seed):
    """The rank's log-probabilities generated on the HOST with the recipe and generator state the reference arm's
    sample uses (cpu_sample_jobs): both arms see identical arrays (the reference arm the first n videos of them)."""
    rng = np.random.default_rng(seed + 3000)
    out = np.empty((int(T.sum()), C), dtype=np.float32)
    pos = 0
    for t, tr in zip(T, trs):
        out[pos:pos + int(t)], _ = synth.planted_logp(rng, int(t), C, tr.tolist(), np.float32)
        pos += int(t)
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (the oracle port of the reference's CPU path) -- checker code, used only here
# ------------------------------------------------------------------------------------------------
_CPU_JOBS = None


def _cpu_decode_ref(i):
    from oracle import hyp_viterbi, poisson
    logp, tr, means = _CPU_JOBS[i]
    t0 = time.perf_counter()
    tab = poisson.poisson_table_loop(means, MAX_LEN)           # PoissonModel(lengths), evaluators.py:167
    hyp_viterbi.decode(logp, [list(map(int, tr))], tab, MAX_LEN, FS)  # vi_decoder.decode, evaluators.py:178
    return logp.shape[0], time.perf_counter() - t0


def _cpu_decode_c(i):
    from oracle import coracle, dense_viterbi, poisson
    logp, tr, means = _CPU_JOBS[i]
    t0 = time.perf_counter()
    rows = coracle.poisson_rows(poisson.poisson_params(means)[tr], FS, MAX_LEN)
    coracle.decode_video(logp, tr, rows, FS, dense_viterbi.numpy_seg0_f32(logp.dtype))
    return logp.shape[0], time.perf_counter() - t0


def cpu_sample_jobs(n_videos, seed=0):
    """The first n videos of the rank-0 split, with host-generated log-probs of the same recipe."""
    T, trs, means = make_split(seed)
    rng = np.random.default_rng(seed + 3000)
    jobs = []
    for v in range(min(n_videos, len(T))):
        lp, _ = synth.planted_logp(rng, int(T[v]), C, trs[v].tolist(), np.float32)
        jobs.append((lp, trs[v], means[v]))
    return jobs


def run_cpu_pool(fn, n_jobs, cores):
    t0 = time.perf_counter()
    if cores == 1:
        res = [fn(i) for i in range(n_jobs)]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(fn, range(n_jobs), chunksize=1)
    wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    return frames / wall, wall, frames, sum(r[1] for r in res)


def cpu_baseline(budget_s=12.0):
    """Times the reference-algorithm port on all host cores over a bounded sample of c2."""
    global _CPU_JOBS
    from oracle import build as oracle_build
    oracle_build.build_oracle()
    cores = os.cpu_count() or 1
    _CPU_JOBS = cpu_sample_jobs(2)
    _, _, f, busy = run_cpu_pool(_cpu_decode_ref, 2, 1)  # calibrate: seconds per frame on one core
    per_frame = busy / f
    n = int(max(cores, min(V_PER_GPU, budget_s * cores / (per_frame * 2300))))
    _CPU_JOBS = cpu_sample_jobs(n)
    fps, wall, frames, _ = run_cpu_pool(_cpu_decode_ref, n, cores)
    fps1 = 1.0 / per_frame
    cfps, cwall, cframes, _ = run_cpu_pool(_cpu_decode_c, n, cores)
    return {
        "value": fps, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"first {n} videos of the c2 split ({frames} frames), oracle/hyp_viterbi.py + loop-built Poisson "
                  f"table per video, {cores} processes, {wall:.1f} s",
        "single_core_value": fps1,
        "c_port": {"value": cfps, "unit": UNIT, "cores": cores,
                   "what": f"oracle/oracle.c dense restatement, same sample, {cwall:.2f} s"},
    }


def main_reference(args, rank, world):
    """The reference arm: the reference's own (CPU, Python) algorithm on the host cores."""
    if rank != 0:
        return
    global _CPU_JOBS
    from oracle import build as oracle_build
    oracle_build.build_oracle()
    cores = os.cpu_count() or 1
    _CPU_JOBS = cpu_sample_jobs(2)
    _, _, f, busy = run_cpu_pool(_cpu_decode_ref, 2, 1)
    per_frame = busy / f
    # each step: a bounded sample, about 6 s of wall time on all cores
    n = int(max(cores, min(V_PER_GPU, 6.0 * cores / (per_frame * 2300))))
    _CPU_JOBS = cpu_sample_jobs(n)
    for _ in range(min(args.warmup, 1)):
        run_cpu_pool(_cpu_decode_ref, min(n, cores), cores)
    walls, frames = [], 0
    for _ in range(args.steps):
        fps, wall, frames, _ = run_cpu_pool(_cpu_decode_ref, n, cores)
        walls.append(wall)
    value = frames * len(walls) / sum(walls)
    sample = (f"first {n} videos of the c2 split ({frames} frames) per step, hypothesis-table port of "
              f"core/viterbi/viterbi.py + loop-built PoissonModel table per video, {cores} processes")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": bench_config(args.gpus, args.collective), "sampled": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from mucon_b200 import dist as mdist
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, ViterbiEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # N > 1: run this rank on the CPUs next to its GPU (NVML's ideal CPU set) BEFORE the pinned host buffers are
    # allocated and first touched, so that the e2e leg's H2D copies read NUMA-local memory
    numa = {"bound": False}
    if world > 1 and not args.no_numa:
        try:
            import pynvml
            pynvml.nvmlInit()
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank), (ncpu + 63) // 64)
            ideal = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1}
            allowed = ideal & set(os.sched_getaffinity(0))
            if allowed and len(allowed) < len(os.sched_getaffinity(0)):
                os.sched_setaffinity(0, allowed)
                numa = {"bound": True, "cpus": len(allowed)}
            else:
                numa = {"bound": False, "why": "the GPU's ideal CPU set does not narrow this process's cpuset",
                        "ideal": len(ideal), "allowed": len(os.sched_getaffinity(0))}
        except Exception as e_:
            numa = {"bound": False, "why": str(e_)[:80]}
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    # Weak scaling: every rank aligns a split of the SAME shape (video lengths, transcripts, length models: seed 0) on
    # its OWN log-probabilities (seed rank), so the per-GPU work really is fixed as N grows -- with a differently seeded
    # split per rank the slowest split's critical path (its longest videos) set the step time (5 % spread at N = 8).
    T, trs, means = make_split(0)
    cands = [[tr.tolist()] for tr in trs]
    host_lp = torch.from_numpy(make_host_logp(T, trs, rank)).pin_memory()
    logp = host_lp.to(device, non_blocking=True)
    torch.cuda.synchronize()
    eng = ViterbiEngine(device)
    params = poisson_params(means)
    payload_cap = 8 * V_PER_GPU + 4 * V_PER_GPU * 12  # same on every rank: [scores | segment lengths | pad]
    plan = AlignPlan(T, cands, C, fs=FS, max_len=MAX_LEN, device=device, len_params=params, labels="best",
                     payload_capacity=payload_cap)
    # N > 1: the kernels store every video's score and segment lengths straight into their slot of every rank's
    # receive buffer (peer-to-peer stores over NVLink, dist.PeerExchange): no collective kernel in the step, the
    # barrier that ends the timed region is the only synchronisation.  --collective nccl keeps the round-1 form
    # (an async all_gather_into_tensor per step, overlapped with the next step).
    pg = px = None
    if world > 1:
        plan_b = AlignPlan(T, cands, C, fs=FS, max_len=MAX_LEN, device=device, len_params=params, labels="best",
                           payload_capacity=payload_cap)
        if args.collective == "nccl":
            pg = mdist.PipelinedGather([plan, plan_b])
        else:
            px = mdist.PeerExchange([plan, plan_b])
    step_no = [0]

    def step(mode="auto"):
        if pg is None and px is None:
            eng.run(plan, logp, seg0_f32=True, mode=mode, write_bs=False)
            return
        i = step_no[0]
        step_no[0] += 1
        if px is not None:
            eng.run(px.plan(i), logp, seg0_f32=True, mode=mode, write_bs=False)
            return
        eng.run(pg.acquire(i), logp, seg0_f32=True, mode=mode, write_bs=False)
        pg.gather(i)

    def barrier():
        if pg is not None:
            pg.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert eng.last_mode == "fused", "the fused kernel did not run"
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    # The whole loop is bracketed, not each step (per-step events add ~3 us of device time each).
    t_all0.record()
    for i in range(args.steps):
        step()
    t_all1.record()
    barrier()
    total_ms = t_all0.elapsed_time(t_all1)
    launches = eng.launches - launches0
    comm = None
    if px is not None:
        # outside the timed region: what the exchange delivered against an NCCL all_gather of the same payloads
        last = step_no[0] - 1
        want = mdist.gather_payload(px.plan(last))
        torch.cuda.synchronize()
        okt = torch.tensor([int(torch.equal(px.result(last), want))], device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        comm = {"kind": "peer-to-peer stores of scores + segment lengths from the alignment kernels' epilogue into "
                        "every rank's receive buffer (CUDA IPC over NVLink); no collective kernel in the step",
                "bytes_per_rank_per_step": int(payload_cap) * world, "equals_nccl_all_gather": bool(okt.item())}
    elif pg is not None:
        comm = {"kind": "NCCL all_gather_into_tensor per step, overlapped with the next step",
                "bytes_per_rank_per_step": int(payload_cap) * world}
    # the alignment launches alone (no collective), same number of steps, for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        eng.run(plan, logp, seg0_f32=True, mode="auto", write_bs=False)
    k1.record()
    barrier()
    # a launch cannot take longer than the step that contains it: when this second loop comes out slower than the
    # timed steps (run-to-run noise of a few us, seen at N > 1), the step time is the better estimate
    fused_ms = min(k0.elapsed_time(k1) / args.steps, total_ms / args.steps)

    # the two-kernel path (what candidate sets use): scan kernel and DP kernel timed separately
    for _ in range(3):
        step("split")
    barrier()
    sev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(5)]
    for e in sev:
        e[0].record()
        eng.run(plan, logp, seg0_f32=True, mode="split", mid_event=e[1])
        e[2].record()
    barrier()
    scan_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in sev]))
    dp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in sev]))

    # ---- full test-time inference of the hot path: backbone forward (2048-d features -> log-probs,
    # tcgen05 projection + conv layers) chained into the fused alignment, all resident in HBM
    full = None
    if not args.no_backbone:
        try:
            from mucon_b200.temporal import MuConBackbone
            torch.manual_seed(rank)
            net = MuConBackbone().eval().to(device)
            bplan = net.plan(T, device)
            feats = torch.empty(int(T.sum()), 2048, device=device)
            feats.normal_(generator=torch.Generator(device).manual_seed(rank)).abs_().mul_(0.5)

            def backbone_step():
                # features -> pooled-resolution log-probabilities (the [T, C] expansion is never written)
                return net.infer_pooled_packed(feats, bplan)

            def full_step():
                lsm, zoff = backbone_step()
                eng.run(plan, lsm, seg0_f32=True, write_bs=False, z_off=zoff)

            def expanded_step():
                # the materialising path (what the drop-in signatures return): [sum T, C] log-probs to HBM and back
                lp = net.logprobs_packed(net.encode_packed(feats, bplan), bplan)
                eng.run(plan, lp, seg0_f32=True, write_bs=False)

            def timed(fn, n):
                for _ in range(2):
                    fn()
                barrier()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b_.record()
                barrier()
                return a.elapsed_time(b_) / n

            nfull = 10
            full_ms = timed(full_step, nfull)
            bb_ms = timed(backbone_step, nfull)
            enc_ms = timed(lambda: net.encode_packed(feats, bplan), nfull)
            exp_ms = timed(expanded_step, nfull)
            w_ = net.ft._weights()
            from mucon_b200 import temporal as _tm
            proj_ms = timed(lambda: _tm.gemm_tf32_bias_act(feats, w_["first_w"], w_["first_b"], True,
                                                           out_dtype=torch.float16), nfull)
            # the s-head on the same batch (BiLSTM encoder over the pooled sequence + attention decoder, teacher-forced
            # with the split's transcripts = the alignment mode of the evaluator): two kernels + three small GEMMs
            shead_ms = None
            try:
                from mucon_b200.shead import SHead
                sh_ = SHead(num_classes=C).eval().to(device)
                _, zoff_, z_ = net.infer_pooled_packed(feats, bplan, want_z=True)
                tf_in_ = [np.concatenate([[C + 1], tr]).astype(np.int32) for tr in trs]
                shead_ms = timed(lambda: sh_.forward_packed(z_, zoff_, bplan.off_host[-1], tf_in_, teacher_forcing=True), 3)
                del sh_, z_
            except Exception as e_:
                shead_ms = "error: " + str(e_)[:120]
            Tsum_ = int(T.sum())
            # backbone reads 4*T*D bytes of features and writes the log-probabilities (here at the
            # pooled resolution: 4*Tz*C); flops = 1.014 MFLOP per frame
            bb_bytes = Tsum_ * 2048 * 4 + Tsum_ * C * 4            # 4*T*D read + 4*T*C log-probs written
            bb_bytes_moved = Tsum_ * 2048 * 4 + int(bplan.rows[-1]) * C * 4   # what this path moves: the pooled table only
            bb_flops = 1.014e6 * Tsum_
            peaks_ = {}
            try:
                peaks_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm_ = float(peaks_.get("hbm_gbs", 6650.0))
            tf_ = float(peaks_.get("bf16_tflops_sustained", 1400.0))
            full = {"what": "backbone forward (TF32 tcgen05 projection from the fp32 features, fp16 tcgen05 WaveNet layers "
                            "with resident weights, fp32 GroupNorm / classifier / log-softmax at the pooled resolution) -> "
                            "fused Viterbi alignment reading the pooled table, 1712 videos/GPU, features resident in HBM",
                    "precision": _tm.DEFAULT_PRECISION,
                    "ms_per_step": full_ms, "backbone_ms": bb_ms, "encode_ms": enc_ms, "projection_ms": proj_ms,
                    "expanded_path_ms_per_step": exp_ms, "shead_ms": shead_ms,
                    "frames_per_sec_per_gpu": float(T.sum()) / (full_ms * 1e-3),
                    "feature_bytes": int(feats.numel() * 4),
                    "feature_read_gbs": feats.numel() * 4 / (bb_ms * 1e-3) / 1e9,
                    "roofline": {"bound": "hbm", "kernel": "backbone forward (projection + 11 layer launches + tail)",
                                 "achieved": bb_bytes / (bb_ms * 1e-3) / 1e9, "peak": hbm_, "unit": "GB/s",
                                 "frac": bb_bytes / (bb_ms * 1e-3) / 1e9 / hbm_, "bytes_per_step": bb_bytes,
                                 "algorithmic_bytes": "4*T*D features + 4*T*C log-probabilities (SURVEY.md 8d)",
                                 "bytes_moved_per_step": bb_bytes_moved,
                                 "note": "the path writes the log-probabilities at the pooled resolution only (4*Tz*C = 1/16 of "
                                         "4*T*C) and the alignment reads them from there; frac_of_moved_bytes counts just that",
                                 "frac_of_moved_bytes": bb_bytes_moved / (bb_ms * 1e-3) / 1e9 / hbm_,
                                 "tensor": {"achieved": bb_flops / (bb_ms * 1e-3) / 1e12, "peak": tf_, "unit": "TFLOP/s",
                                            "frac": bb_flops / (bb_ms * 1e-3) / 1e12 / tf_,
                                            "flops": "1.014 MFLOP per frame (SURVEY.md 8d)",
                                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (the projection "
                                                           "runs in TF32, half that rate)"}}}
            del feats, net
            torch.cuda.empty_cache()
        except Exception as e:  # the headline number must not depend on this extra leg
            full = {"error": str(e)[:200]}

    # ---- mask generation of the consistency loss for the same split (forward + backward) ----------
    masks_leg = None
    try:
        from mucon_b200.masks import batch_meta, create_masks_batch
        Ms = [len(t) for t in trs]
        Tl = [int(t) for t in T]
        rngm = np.random.default_rng(1000 + rank)
        Lh = np.concatenate([float(t) * rngm.dirichlet(3 * np.ones(m)) for t, m in zip(T, Ms)]).astype(np.float32)
        Ld = torch.from_numpy(Lh).to(device).requires_grad_(True)
        mask_bytes = int(sum(4 * t * m for t, m in zip(Tl, Ms)))
        meta = batch_meta(Tl, Ms, device)   # offset tables built once, like an AlignPlan
        for _ in range(3):
            mo, _off = create_masks_batch(Tl, Ld, Ms, meta=meta)
        go = torch.ones_like(mo)
        for _ in range(3):
            Ld.grad = None
            mo.backward(go, retain_graph=True)
        barrier()
        # through the C ABI with preallocated buffers (the autograd entry point above adds ~60 us of host
        # time per call, more than the kernels take)
        from mucon_b200 import masks as mmod
        n_off_d, T_d, off_d, rv_d, Vm, n_rows, max_Tm, total = meta[0]
        Lc = Ld.detach().clone()
        ws = torch.empty(2 * n_rows, dtype=torch.float32, device=device)
        gL = torch.empty(n_rows, dtype=torch.float32, device=device)
        mo = mo.detach()
        for _ in range(3):
            mmod._launch_fwd(Lc, n_off_d, T_d, off_d, rv_d, Vm, n_rows, max_Tm, 0.0, 0, 0, None, mo)
            mmod._launch_bwd(Lc, n_off_d, T_d, off_d, rv_d, Vm, n_rows, 0.0, 0, 0, go, ws, gL)
        barrier()
        m0, m1, m2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        nm = 20
        m0.record()
        for _ in range(nm):
            mmod._launch_fwd(Lc, n_off_d, T_d, off_d, rv_d, Vm, n_rows, max_Tm, 0.0, 0, 0, None, mo)
        m1.record()
        for _ in range(nm):
            mmod._launch_bwd(Lc, n_off_d, T_d, off_d, rv_d, Vm, n_rows, 0.0, 0, 0, go, ws, gL)
        m2.record()
        barrier()
        fwd_ms, bwd_ms = m0.elapsed_time(m1) / nm, m1.elapsed_time(m2) / nm
        masks_leg = {"what": "create_masks for all 1712 videos in one launch (box template), %d mask rows" % sum(Ms),
                     "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "bytes_written_fwd": mask_bytes,
                     "fwd_gbs": mask_bytes / (fwd_ms * 1e-3) / 1e9, "bwd_read_gbs": mask_bytes / (bwd_ms * 1e-3) / 1e9,
                     "note": "bound: HBM write (fwd) / read (bwd) of 4*N*T bytes (110 MB: fits the 126 MB L2, so "
                     "the backward's reads are L2 hits); back-to-back launches through the C ABI"}
        # fused flint evidence E = masks @ logits without the masks (reads 4*T*C bytes per video)
        try:
            from mucon_b200.loss import _flint_meta, flint_evidence
            fmeta = _flint_meta(Ms, Tl, device)
            segd = logp.detach().clone().requires_grad_(True)
            Lf = Lc.clone().requires_grad_(True)
            for _ in range(3):
                Ev = flint_evidence(Lf, segd, Ms, Tl, meta=fmeta)
                Ev.backward(torch.ones_like(Ev))
                segd.grad = None
                Lf.grad = None
            barrier()
            q0, q1, q2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            nq = 10
            q0.record()
            for _ in range(nq):
                Ev = flint_evidence(Lf, segd, Ms, Tl, meta=fmeta)
            q1.record()
            gEv = torch.ones_like(Ev)
            for _ in range(nq):
                segd.grad = None
                Lf.grad = None
                Ev.backward(gEv, retain_graph=True)
            q2.record()
            barrier()
            ff, fb = q0.elapsed_time(q1) / nq, q1.elapsed_time(q2) / nq
            seg_bytes = int(logp.numel() * 4)
            masks_leg["flint_fused"] = {"what": "E = masks @ frame logits for all 1712 videos, masks never written",
                                        "fwd_ms": ff, "fwd_read_gbs": seg_bytes / (ff * 1e-3) / 1e9, "bwd_ms": fb,
                                        "bwd_note": "grad wrt the frame logits (740 MB written) and the lengths"}
            del segd, Lf, Ev, gEv
        except Exception as e:
            masks_leg["flint_fused"] = {"error": str(e)[:200]}
        del mo, go, Ld
        torch.cuda.empty_cache()
    except Exception as e:
        masks_leg = {"error": str(e)[:200]}

    # ---- e2e: host API, pinned host log-probs in, labels + scores + segments out ---------------
    host_logp = host_lp   # the pinned host array both arms are built from
    n_lab = plan.n_labels
    host_labels = torch.empty(n_lab, dtype=torch.int32, pin_memory=True)
    host_small = torch.empty(plan.U, dtype=torch.float64, pin_memory=True)
    host_seg = torch.empty(int(plan.tr_off[-1]), dtype=torch.int32, pin_memory=True)
    dev_in = torch.empty_like(logp)

    def e2e_step():
        dev_in.copy_(host_logp, non_blocking=True)   # the 740 MB copy runs while the host builds the plan
        p = AlignPlan(T, cands, C, fs=FS, max_len=MAX_LEN, device=device, len_params=poisson_params(means),
                      labels="best")
        eng.run(p, dev_in, seg0_f32=True, write_bs=False)
        host_labels.copy_(p.labels, non_blocking=True)
        host_small.copy_(p.score, non_blocking=True)
        host_seg.copy_(p.seg_blocks, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return p

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        p_last = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        p_last = e2e_step()
    barrier()
    e2e_serial_s = (time.perf_counter() - t0) / e2e_steps
    # The same steps through the streaming API (viterbi.HostAlignPipeline): every step still copies its 740 MB of
    # log-probabilities host -> device and its labels / scores / segment lengths device -> host, but step i's kernels and
    # device -> host copy run under step i + 1's host -> device copy (three streams, two device input buffers).
    from mucon_b200.viterbi import HostAlignPipeline
    pipe = HostAlignPipeline(eng)

    def mk_plan():
        return AlignPlan(T, cands, C, fs=FS, max_len=MAX_LEN, device=device, len_params=poisson_params(means), labels="best")

    def e2e_stream(n):
        prev, last = None, None
        for _ in range(n):
            tk = pipe.submit(host_logp, mk_plan, seg0_f32=True)
            if prev is not None:
                last = pipe.result(prev)
            prev = tk
        return pipe.result(prev)

    e2e_stream(3)
    barrier()
    t0 = time.perf_counter()
    res_last = e2e_stream(e2e_steps)
    barrier()
    e2e_stream_s = (time.perf_counter() - t0) / e2e_steps
    # both are user-facing calls of the package; the line reports the faster one and both times (with several ranks on
    # one host the streamed form's concurrent H2D + D2H + plan builds contend on the host side and the serial step wins)
    e2e_s = min(e2e_stream_s, e2e_serial_s)
    e2e_mode = "streamed (viterbi.HostAlignPipeline)" if e2e_stream_s <= e2e_serial_s else "serial (AlignPlan + ViterbiEngine.run per step)"
    e2e_same = bool(torch.equal(res_last[1], host_labels) and torch.equal(res_last[2], host_small) and
                    torch.equal(res_last[3], host_seg))   # the streamed results equal the serial step's
    # where an e2e step goes: the parts timed one at a time (in the step itself the plan build overlaps the H2D copy)
    def wall(fn, n=3):
        torch.cuda.synchronize()
        t_ = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t_) / n * 1e3
    plan_ms = wall(lambda: AlignPlan(T, cands, C, fs=FS, max_len=MAX_LEN, device=device,
                                     len_params=poisson_params(means), labels="best"))
    h2d_ms = wall(lambda: dev_in.copy_(host_logp, non_blocking=True))
    kern_ms = wall(lambda: eng.run(p_last, dev_in, seg0_f32=True, write_bs=False))
    d2h_ms = wall(lambda: (host_labels.copy_(p_last.labels, non_blocking=True),
                           host_small.copy_(p_last.score, non_blocking=True),
                           host_seg.copy_(p_last.seg_blocks, non_blocking=True)))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    h2d = logp.numel() * logp.element_size() + p_last.h2d_meta_bytes
    d2h = host_labels.numel() * 4 + host_small.numel() * 8 + host_seg.numel() * 4

    # ---- the other configs (bench_legs.py) ---------------------------------------------------------
    legs = {}
    if not args.no_legs:
        import bench_legs
        del dev_in
        torch.cuda.empty_cache()
        try:
            legs["c3_candidates"] = bench_legs.leg_c3(device, rank, world, make_split, device_logp)
        except Exception as e:
            legs["c3_candidates"] = {"error": str(e)[:300]}
            if world > 1:
                raise
        if rank == 0:
            for name, fn in (("c1_single_video", lambda: bench_legs.leg_c1(device)),
                             ("c4_long_video", lambda: bench_legs.leg_c4(device)),
                             ("length_distributions", lambda: bench_legs.leg_distributions(device)),
                             ("train_step", lambda: bench_legs.leg_train(device, make_split))):
                try:
                    legs[name] = fn()
                except Exception as e:
                    legs[name] = {"error": str(e)[:300]}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # max over ranks
    per_rank_ms = [total_ms / args.steps]
    per_rank_e2e = [e2e_s * 1e3]
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, e2e_serial_s * 1e3, e2e_stream_s * 1e3], dtype=torch.float64, device=device)
        allr = torch.empty(3 * world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, 3)
        per_rank_ms = allr[:, 0].tolist()
        # one mode for the whole job: the one whose slowest rank is faster
        e2e_serial_s, e2e_stream_s = float(allr[:, 1].max()) * 1e-3, float(allr[:, 2].max()) * 1e-3
        streamed = e2e_stream_s <= e2e_serial_s
        e2e_s = min(e2e_stream_s, e2e_serial_s)
        e2e_mode = "streamed (viterbi.HostAlignPipeline)" if streamed else "serial (AlignPlan + ViterbiEngine.run per step)"
        per_rank_e2e = allr[:, 2 if streamed else 1].tolist()
        t = torch.tensor([total_ms, scan_ms, dp_ms, fused_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, scan_ms, dp_ms, fused_ms = t.tolist()
        fr = torch.tensor([plan.aligned_frames], dtype=torch.float64, device=device)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
        frames_all = float(fr.item())
    else:
        frames_all = float(plan.aligned_frames)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        ms_per_step = total_ms / args.steps
        # algorithmic bytes: scan reads 4*T*C per video; DP unit writes 4T labels,
        # 2KN back-pointer bytes, 528N Poisson rows, 8+8N score/segments
        Tsum, Ksum, Nsum = int(T.sum()), int((T // FS).sum()), int(sum(len(t) for t in trs))
        KN = int(((T // FS) * np.array([len(t) for t in trs])).sum())
        scan_bytes = 4 * Tsum * C
        dp_bytes = 4 * Tsum + 2 * KN + 528 * Nsum + 8 * len(T) + 8 * Nsum
        path_bytes = scan_bytes + dp_bytes
        fused_gbs = path_bytes / (fused_ms * 1e-3) / 1e9
        scan_gbs = scan_bytes / (scan_ms * 1e-3) / 1e9
        traffic, traffic_src = (None, None) if (args.no_traffic or world > 1) else measure_traffic()
        if traffic is None:
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("align_fused_kernel_dram_bytes")
                traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum of both launches from the ncu --set full capture "
                               "profiles/r1d_align_fused_pair_c2.md (profiles/traffic.json); not re-measured in this run" +
                               (" (N > 1 or --no-traffic)" if (args.no_traffic or world > 1) else " (ncu did not run: %s)" % traffic_src))
            except Exception:
                pass
        out = {
            "metric": METRIC, "value": frames_all / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic",
            "config": bench_config(world, args.collective),
            "clocks": sampler.summary(),
            "e2e": {"value": frames_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "what": "per step AlignPlan build + pinned H2D of the log-probs + fused kernel + D2H of labels/scores/segments; "
                            "streamed = viterbi.HostAlignPipeline (step i's kernels and D2H under step i+1's H2D, three "
                            "streams), serial = one step at a time; the faster of the two is the value",
                    "mode": e2e_mode, "serial_ms_per_step": e2e_serial_s * 1e3, "streamed_ms_per_step": e2e_stream_s * 1e3,
                    "streamed_equals_serial": e2e_same,
                    "breakdown_rank0_ms": {"plan_build": plan_ms, "h2d": h2d_ms, "kernel": kern_ms, "d2h": d2h_ms,
                                           "note": "timed one at a time; in the step the plan build overlaps the H2D copy"},
                    "h2d_gbs_rank0": h2d / (h2d_ms * 1e-3) / 1e9,
                    "per_rank_ms": per_rank_e2e, "numa_rank0": numa,
                    "limiter": "host-to-device copy of the log-probabilities (740 MB per rank per step over PCIe); with "
                               "N ranks copying at once the host side is shared, so e2e scales worse than the kernels"},
            "per_rank_ms_per_step": per_rank_ms,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "align_fused_kernel (scan + DP + traceback + labels)" + (
                             ", %d concurrent launches: the %d longest videos with a warp per segment, the rest "
                             "with 8 lanes per segment" % (2, plan.n_long) if plan.n_long else ", one launch"),
                         "achieved": fused_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fused_gbs / hbm_peak,
                         "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peak_src, "bytes_per_launch": path_bytes,
                         "ms_per_launch": fused_ms, "launches_per_step": 2 if plan.n_long else 1,
                         "timing": "CUDA events around %d back-to-back steps on the launching stream / %d" % (args.steps, args.steps),
                         "algorithmic_bytes": "4TC + 4T + 2KN + 528N + 8 + 8N per unit (SURVEY.md 8d)"},
            "two_kernel_path": {"what": "same batch through mucon_viterbi_blockscores + mucon_viterbi_decode "
                                        "(the path candidate sets use)",
                                "scan_kernel": {"achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s",
                                                "frac": scan_gbs / hbm_peak, "bytes_per_launch": scan_bytes,
                                                "ms_per_launch": scan_ms},
                                "dp_kernel_ms": dp_ms},
        }
        if comm is not None:
            out["comm"] = comm
        if full is not None:
            out["full_inference"] = full
        if masks_leg is not None:
            if "fwd_ms" in masks_leg:
                masks_leg["roofline"] = {"bound": "hbm", "kernel": "masks_fwd_kernel", "achieved": masks_leg["fwd_gbs"],
                                         "peak": hbm_peak, "unit": "GB/s", "frac": masks_leg["fwd_gbs"] / hbm_peak,
                                         "bytes_per_launch": masks_leg["bytes_written_fwd"],
                                         "algorithmic_bytes": "4*N*T written per video (SURVEY.md 8d)"}
                ff_ = masks_leg.get("flint_fused", {})
                if "fwd_read_gbs" in ff_:
                    ff_["roofline"] = {"bound": "hbm", "kernel": "flint_fwd_kernel", "achieved": ff_["fwd_read_gbs"],
                                       "peak": hbm_peak, "unit": "GB/s", "frac": ff_["fwd_read_gbs"] / hbm_peak,
                                       "algorithmic_bytes": "4*T*C read + 4*N*C written per video (SURVEY.md 8d)"}
            out["masks"] = masks_leg
        out.update(legs)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline()
            if not args.no_legs:
                import bench_legs
                for name, fn in (("backbone", bench_legs.cpu_backbone_baseline), ("masks", bench_legs.cpu_masks_baseline)):
                    try:
                        out["cpu_baseline"][name] = fn()
                    except Exception as e:
                        out["cpu_baseline"][name] = {"error": str(e)[:200]}
        print(json.dumps(out))
    if px is not None:
        px.close()
    if world > 1:
        dist.destroy_process_group()


def measure_traffic(timeout=240):
    """DRAM bytes of one step's alignment launches, measured NOW on this GPU: ncu (two DRAM counters only) around a child
    process that runs the same plan on the same split (scripts/prof_fused.py), after this process has finished timing.
    Returns (bytes, source) or (None, reason)."""
    import shutil
    import subprocess
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "no ncu"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:align_fused", "-s", "4", "-c", "2", "--csv", sys.executable, os.path.join(ROOT, "scripts", "prof_fused.py")]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except Exception as e:
        return None, type(e).__name__
    import csv
    import io
    rows = [x for x in csv.reader(io.StringIO(r.stdout)) if len(x) > 5]
    hdr = next((x for x in rows if "Metric Value" in x and "Metric Unit" in x), None)
    if hdr is None:
        return None, "no ncu table (rc %d)" % r.returncode
    vi, ui, ki = hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, kernels = 0.0, set()
    for x in rows:
        if x is hdr or x[ui] not in scale:
            continue
        total += float(x[vi].replace(",", "")) * scale[x[ui]]
        kernels.add(x[ki][:40])
    if not total:
        return None, "empty ncu table"
    return int(total), ("measured in this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum around "
                        "scripts/prof_fused.py (same split and plan), both launches of one step: %s" % sorted(kernels))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-backbone", action="store_true", help="skip the backbone + alignment leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the c1 / c3 / c4 / training-step legs")
    ap.add_argument("--no-traffic", action="store_true", help="do not re-measure roofline.traffic with ncu (N = 1 only)")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind ranks to the CPUs next to their GPU")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: result exchange by peer stores from the kernels (default) or an NCCL all_gather per step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
