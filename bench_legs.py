"""Extra legs of bench.py: the other BASELINE.json configs (c1 latency, c3 candidate sets sharded over the ranks,
c4 long video, c5 training step) and the CPU baselines of BASELINE.md section 4 items 2-3 (backbone, masks).
Every leg returns a dict for the JSON line and must never take the headline number down with it (callers wrap
them in try/except).  The oracle is imported only inside the cpu_* functions (checker code, CPU baseline legs)."""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p.get("hbm_gbs", 6650.0)), float(p.get("bf16_tflops_sustained", 1400.0)), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def timed(torch, fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ------------------------------------------------------------------------------------------------------------------
def leg_c1(device):
    """configs[0]: one Breakfast-shaped video (T = 2000, 48 classes, 6 segments) through the reference's own call
    pattern (evaluators.py:147-180: new grammar + PoissonModel + decode per video), host arrays in, lists out."""
    import torch
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, Viterbi, ViterbiEngine
    from tests import synth
    rng = np.random.default_rng(0)
    tr = [0, 5, 7, 5, 12, 0]
    lp, _ = synth.planted_logp(rng, 2000, 48, tr, np.float32)
    means = synth.class_means(rng.dirichlet(5 * np.ones(6)).astype(np.float32), tr, 48, 2000)
    dec = Viterbi(SingleTranscriptGrammar(tr, 48), PoissonModel(means), frame_sampling=30, device=device)
    for _ in range(5):
        dec.decode(lp)
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        dec.grammar = SingleTranscriptGrammar(tr, 48)
        dec.length_model = PoissonModel(means)
        dec.decode(lp)
    wall_us = (time.perf_counter() - t0) / n * 1e6
    eng = ViterbiEngine(device)
    plan = AlignPlan([2000], [[tr]], 48, device=device, len_params=poisson_params(means)[None])
    dlp = torch.from_numpy(lp).to(device)
    k_us = timed(torch, lambda: eng.run(plan, dlp, seg0_f32=True, write_bs=False), 20) * 1e3
    return {"what": "c1: T=2000, C=48, N=6, one video through the reference call pattern (new SingleTranscriptGrammar + "
                    "PoissonModel + Viterbi.decode per video; host float32 array in, score / labels / segments out)",
            "decode_wall_us": wall_us, "kernel_us": k_us, "aligned_frames_per_s": 2000 / (wall_us * 1e-6)}


def leg_c4(device):
    """configs[3]: T = 40 000, 100 classes, 60 segments -- a single serial DP of 1333 steps x 3960 states."""
    import torch
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, ViterbiEngine
    from tests import synth
    r4 = np.random.default_rng(4)
    tr4 = r4.permutation(100)[:60].tolist()
    lp4, _ = synth.planted_logp(r4, 40000, 100, tr4, np.float32)
    m4 = synth.class_means(r4.dirichlet(5 * np.ones(60)).astype(np.float32), tr4, 100, 40000)
    plan4 = AlignPlan([40000], [[tr4]], 100, device=device, len_params=poisson_params(m4)[None])
    d4 = torch.from_numpy(lp4).to(device)
    eng = ViterbiEngine(device)
    ms = timed(torch, lambda: eng.run(plan4, d4, seg0_f32=True, write_bs=False), 5)
    return {"what": "c4: one video, T=40000, C=100, N=60 (kernel time, log-probs resident)", "ms": ms,
            "mode": eng.last_mode, "aligned_frames_per_s": 40000 / (ms * 1e-3)}


def c3_candidates(T, trs, videos, n_cands=64, C=48, fs=30):
    """Candidate transcripts of the given videos; seeded per video, so every rank builds the same set for a video."""
    from tests import synth
    cands, means = [], []
    for v in videos:
        r = np.random.default_rng(50000 + int(v))
        K = int(T[v]) // fs
        cands.append(synth.random_edits(r, trs[v], C, n_cands, max(2, -(-K // 66)), min(30, K)))
        means.append(synth.class_means(r.dirichlet(np.ones(len(trs[v]))).astype(np.float32), trs[v], C, int(T[v])))
    return cands, np.stack(means)


def leg_c3(device, rank, world, make_split, device_logp, steps=5):
    """configs[2]: every video of the c2 split (seed 0) with 64 candidate transcripts, the videos sharded over the
    ranks (dist.shard_videos: greedy by T x candidates, all candidates of a video on one rank) -- STRONG scaling.
    Per step and rank: block-score scan + lane-per-segment DP of its units + per-video arg-max + labels of the
    winners; all units' scores / segment lengths land on every rank (dist.PeerExchange when world > 1)."""
    import torch
    import torch.distributed as dist
    from mucon_b200 import dist as mdist
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, ViterbiEngine
    T, trs, _ = make_split(0)
    NC = 64
    mine = mdist.shard_videos(T, [NC] * len(T), world)[rank]
    cands, means = c3_candidates(T, trs, mine, NC)
    Tm = T[mine]
    logp = device_logp(Tm, [trs[v] for v in mine], 7000 + rank, device)
    params = poisson_params(means)
    n_pos = sum(len(t) for cl in cands for t in cl)
    need = torch.tensor([8 * NC * len(mine) + 4 * n_pos], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
    cap = int(need.item())
    from mucon_b200.viterbi import FlatCandidates
    t0 = time.perf_counter()
    flat = FlatCandidates.from_lists(cands)   # candidate sets as arrays (how a beam search would hand them over) + the
    flat_s = time.perf_counter() - t0         # reference's tie order between candidates (grammar.tie_ranks, Python sets)
    AlignPlan(Tm[:8], FlatCandidates.from_lists(cands[:8]), 48, device=device, len_params=params[:8], labels="best")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = AlignPlan(Tm, flat, 48, device=device, len_params=params, labels="best", payload_capacity=cap)
    torch.cuda.synchronize()
    prep_s = time.perf_counter() - t0
    eng = ViterbiEngine(device)
    px = mdist.PeerExchange([plan]) if world > 1 else None

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(2):
        eng.run(plan, logp, seg0_f32=True)
    fence()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        eng.run(plan, logp, seg0_f32=True)
    b.record()
    fence()
    ms = a.elapsed_time(b) / steps
    frames = float(plan.aligned_frames)
    ok = None
    if world > 1:
        t = torch.tensor([ms, prep_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, prep_s = t.tolist()
        fr = torch.tensor([frames], dtype=torch.float64, device=device)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
        frames = float(fr.item())
        want = mdist.gather_payload(plan)   # outside the timed region: the exchange against an NCCL all_gather
        torch.cuda.synchronize()
        okt = torch.tensor([int(torch.equal(px.result(0), want))], device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
        px.close()
    return {"what": "c3: 1712 videos x 64 candidate transcripts (109568 units), videos sharded over the ranks "
                    "(strong scaling): scan + DP of every unit + per-video arg-max + winner labels",
            "n_gpus": world, "scaling": "strong", "mode": eng.last_mode, "units_this_rank": int(plan.U),
            "ms_per_step": ms, "aligned_frames_per_s": frames / (ms * 1e-3), "plan_build_s_max_over_ranks": prep_s,
            "lists_to_arrays_and_tie_ranks_s_rank0": flat_s,
            "exchange": ("peer stores from the kernel epilogue, equal to an NCCL all_gather: %s" % ok) if world > 1 else "none"}


def leg_train(device, make_split):
    """configs[4]: training step -- backbone forward + GroupNorm / classifier tail + batched flint consistency loss
    (mask generation fused into the evidence kernels) + backward + SGD, 32 videos with c2's T distribution, one
    CUDA-graph replay per step (mucon_b200.train.TrainStep)."""
    import torch
    from mucon_b200 import train
    from mucon_b200.temporal import MuConBackbone
    T_all, _, _ = make_split(0)
    rng = np.random.default_rng(5)
    Ts = [int(t) for t in T_all[:32]]
    Ns = [int(rng.integers(2, 13)) for _ in Ts]
    torch.manual_seed(0)
    m = MuConBackbone().to(device).train()          # dropout 0.25 as configured (default.py:81-96)
    opt = torch.optim.SGD(m.parameters(), lr=1e-3)
    ts = train.TrainStep(m, Ts, Ns, optimizer=opt)
    ts.feats.copy_(torch.randn(ts.feats.shape, device=device).abs() * 0.5)
    ts.transcripts.copy_(torch.from_numpy(np.concatenate([rng.integers(0, 48, n) for n in Ns])).to(device))
    ms = timed(torch, ts.run, 20, warm=2)
    eager = train.TrainStep(m, Ts, Ns, optimizer=opt, graph=False)
    eager.feats.copy_(ts.feats)
    eager.transcripts.copy_(ts.transcripts)
    ms_eager = timed(torch, eager.run, 5, warm=2)
    frames = int(sum(Ts))
    flops = 3 * 1.014e6 * frames
    _, tf_peak, src = peaks()
    hbm, _, _ = peaks()
    bytes_ = 2 * frames * 2048 * 4      # the features are read by the projection and again by its weight gradient
    return {"what": "c5: 32 videos (T from the c2 distribution), 2048-d features, forward + tail + flint consistency "
                    "loss + backward + SGD step, dropout 0.25, one CUDA-graph replay per step",
            "dtype": "tf32 operands (tcgen05 kind::tf32), fp32 accumulation / activations / gradients",
            "videos": 32, "frames": frames, "ms_per_step": ms, "ms_per_step_eager": ms_eager,
            "frames_per_s": frames / (ms * 1e-3), "videos_per_s": 32 / (ms * 1e-3),
            "tflops": flops / (ms * 1e-3) / 1e12, "flops_per_step": "3 x 1.014 MFLOP/frame (fwd + dgrad + wgrad, SURVEY.md 8d)",
            "roofline": {"bound": "launch latency at this batch size (about 200 kernels of 5-60 us); floors: "
                                  "tensor %.2f ms at half of %s bf16 = TF32 rate, HBM %.2f ms" % (
                                      flops / (0.5 * tf_peak * 1e12) * 1e3, src, bytes_ / (hbm * 1e9) * 1e3),
                         "tensor_frac": flops / (ms * 1e-3) / 1e12 / (0.5 * tf_peak)}}


# ------------------------------------------------------------------------------------------------------------------
# CPU baselines of the other two parts of the path: the reference modules restated
# under oracle/ (pinned to the reference by tests/golden/backbone.npz, masks.npz, loss.npz), on the host cores.
def cpu_backbone_baseline(budget_s=8.0):
    import torch
    import torch.nn.functional as F  # noqa: F401
    from oracle import backbone as obb
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(0)
    m = MuConBackbone().eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    T = 2000
    feats = torch.randn(1, T, 2048).abs() * 0.5
    with torch.no_grad():
        for _ in range(2):
            obb.logprobs(sd, obb.encode(sd, feats, m.ft.stages, m.ft.pooling_layers), T)
        n, t0 = 0, time.perf_counter()
        while True:
            obb.logprobs(sd, obb.encode(sd, feats, m.ft.stages, m.ft.pooling_layers), T)
            n += 1
            if n >= 10 and time.perf_counter() - t0 > budget_s / 2 or time.perf_counter() - t0 > budget_s:
                break
        dt = (time.perf_counter() - t0) / n
    return {"value": T / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d x one c1 video (T=2000, D=2048): oracle/backbone.py restatement of WaveNetBlock + GroupNorm + "
                      "ReLU + interpolate + classifier + log_softmax, torch fp32 eval, no_grad, %.1f ms per video" % (n, dt * 1e3)}


def cpu_masks_baseline():
    import torch
    from oracle import masks as omasks
    rng = np.random.default_rng(0)
    T, N = 2000, 6
    L = torch.from_numpy((T * rng.dirichlet(3 * np.ones(N))).astype(np.float32))
    for _ in range(5):
        omasks.create_masks_torch(T, L.clone(), 0.0, "box", align_corners=False)
    n, t0 = 50, time.perf_counter()
    for _ in range(n):
        omasks.create_masks_torch(T, L.clone(), 0.0, "box", align_corners=False)
    dt = (time.perf_counter() - t0) / n
    return {"value": 4.0 * N * T / dt / 1e9, "unit": "GB/s of mask output", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "50 x create_masks(T=2000, N=6, box): oracle/masks.py restatement (cumsum -> affine_grid -> "
                      "grid_sample), %.3f ms per call" % (dt * 1e3)}


def leg_distributions(device, steps=20):
    """The long-tail launch policy on length distributions it was NOT tuned on: the alignment of 1712 videos with equal
    lengths, a bimodal mix and c2's log-normal, each as one launch (long_K = 0) and with the automatic policy."""
    import torch
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, ViterbiEngine
    from tests import synth
    eng = ViterbiEngine(device)
    res = {}
    V, C = 1712, 48
    for name in ("lognormal_c2", "equal_2250", "bimodal_650_9000"):
        rng = np.random.default_rng(23)
        if name == "lognormal_c2":
            T, trs = synth.breakfast_split(seed=0, V=V, C=C, fs=30, J=66)
        else:
            T = np.full(V, 2250) if name.startswith("equal") else \
                np.where(rng.random(V) < 0.75, rng.integers(500, 800, V), rng.integers(8500, 9500, V))
            trs = []
            for t in T:
                K = int(t) // 30
                trs.append(rng.integers(0, C, int(rng.integers(max(2, -(-K // 66)), min(12, K) + 1))))
        means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, C, int(t))
                          for tr, t in zip(trs, T)])
        logp = torch.log_softmax(torch.randn(int(T.sum()), C, device=device), dim=1).contiguous()
        row = {"frames": int(T.sum())}
        for pol, long_K in (("single_launch", 0), ("auto", None)):
            plan = AlignPlan(T, [[tr.tolist()] for tr in trs], C, device=device, len_params=poisson_params(means), long_K=long_K)
            ms = timed(torch, lambda: eng.run(plan, logp, seg0_f32=True, write_bs=False), steps)
            bytes_ = 4 * int(T.sum()) * C + 4 * int(T.sum())
            row[pol] = {"ms": ms, "n_long": int(plan.n_long), "gbs": bytes_ / (ms * 1e-3) / 1e9}
        res[name] = row
        del logp
    hbm, _, _ = peaks()
    for r in res.values():
        for pol in ("single_launch", "auto"):
            r[pol]["frac_of_hbm_peak"] = r[pol]["gbs"] / hbm
    res["what"] = ("fused alignment of 1712 videos per length distribution: one launch vs the automatic long-tail policy "
                   "(videos within 15 % of the longest, at most 40, get a wide launch of their own); GB/s counts 4TC + 4T")
    return res
