"""GPU parity: the CUDA Viterbi path (through the C ABI) against the oracle and the frozen
reference outputs.  Integer outputs (labels, segments, back-pointers, final j) and block scores
must be bit-exact; the float64 path score must be equal to the last bit."""
import os

import numpy as np
import pytest
import torch

from oracle import coracle, dense_viterbi, hyp_viterbi, poisson
from tests import synth
from tests.util import golden_names, load_golden, same_score

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(cuda_device):
    from mucon_b200.viterbi import ViterbiEngine
    return ViterbiEngine(cuda_device)


@pytest.fixture(params=["auto", "split", "lanes"])
def mode(request):
    """auto = fused single-launch kernel where it applies (falls back to split); split = scan + DP kernels."""
    return request.param


def run_units(eng, logps, cands, means, fs=30, max_len=2000, seg0=None, labels="all", use_rows=False, mode="auto",
              force_generic=False):
    """logps: list of [T,C] arrays (one per video); cands: per video list of transcripts; means: per video [C]."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan
    C = logps[0].shape[1]
    kw = {}
    if use_rows:
        J = max_len // fs
        kw["len_rows"] = [coracle.poisson_rows(poisson.poisson_params(m)[tr], fs, max_len)
                          for m, cl in zip(means, cands) for tr in cl]
    else:
        kw["len_params"] = np.stack([poisson_params(m) for m in means])
    plan = AlignPlan([l.shape[0] for l in logps], cands, C, fs=fs, max_len=max_len, device=eng.device,
                     labels=labels, force_generic=force_generic, **kw)
    packed = torch.from_numpy(np.concatenate(logps)).to(eng.device)
    if mode == "lanes" and plan.lane_unit is None:
        pytest.skip("lane-per-segment kernel covers J <= 66, N <= 33")
    eng.run(plan, packed, seg0_f32=seg0, mode=mode)
    torch.cuda.synchronize()
    return plan, eng.fetch(plan, want_bp=True)


def oracle_unit(logp, tr, means, fs, max_len, seg0):
    rows = coracle.poisson_rows(poisson.poisson_params(means)[tr], fs, max_len)
    bs = coracle.block_scores(logp, fs)
    d = coracle.viterbi(bs, tr, rows, seg0)
    d["labels"] = coracle.decode_video(logp, tr, rows, fs, seg0)["labels"]
    d["bs"] = bs
    return d


def check_unit(plan, out, u, ref, T):
    a, b = plan.tr_off[u], plan.tr_off[u + 1]
    assert out["status"][u] in (0, 2)
    assert same_score(out["score"][u], ref["score"]), (out["score"][u], ref["score"])
    assert np.array_equal(out["seg_blocks"][a:b], ref["seg_blocks"])
    if out["status"][u] == 0:
        assert out["final_j"][u] == ref["jf"]
        K, N = ref["bp"].shape
        bp = out["bp"][plan.bp_off[u]:plan.bp_off[u + 1]].reshape(K, N)
        assert np.array_equal(bp.astype(np.uint16), ref["bp"].astype(np.uint16))
    lo = plan.lab_off[u]
    assert np.array_equal(out["labels"][lo:lo + T], ref["labels"])


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixture(eng, name, mode):
    """Frozen outputs of the unmodified reference decoder (tests/golden/make_golden.py)."""
    g = load_golden(name)
    T = g["logp"].shape[0]
    plan, out = run_units(eng, [g["logp"]], [g["transcripts"]], [g["means"]], g["fs"], g["max_len"],
                          seg0=g["seg0_f32"], labels="best", mode=mode)
    if plan.single:
        u = 0
    else:
        u = int(out["best"][0])
    assert same_score(out["score"][u], g["score"])
    assert np.array_equal(out["labels"][:T], g["labels"])
    tr = g["transcripts"][u]
    sb = out["seg_blocks"][plan.tr_off[u]:plan.tr_off[u + 1]]
    assert dense_viterbi.segments_from_blocks(sb, tr, g["fs"], T) == g["segments"]
    if plan.single and np.isfinite(g["score"]):
        K = T // g["fs"]
        assert np.array_equal(out["bp"][:K * len(tr)].reshape(K, len(tr)).astype(np.uint16), g["bp"].astype(np.uint16))
    if name.startswith("generic"):
        assert eng.last_mode == "generic"


@pytest.mark.parametrize("dtype,seg0", [(np.float32, True), (np.float32, False), (np.float64, False)])
def test_random_batch_bit_exact(eng, dtype, seg0, mode):
    rng = np.random.default_rng(42)
    logps, cands, means = [], [], []
    for i in range(24):
        C = 48
        N = int(rng.integers(1, 13))
        K = int(rng.integers(max(N, 1), min(N * 66, 340) + 1))
        T = K * 30 + int(rng.integers(0, 30))
        tr = rng.integers(0, C, N).tolist()
        lp, _ = synth.planted_logp(rng, T, C, tr, dtype)
        if i % 5 == 0:
            lp = np.round(lp)  # integer-valued: exercises ties
        logps.append(lp)
        cands.append([tr])
        means.append(synth.class_means(rng.dirichlet(np.ones(N)).astype(np.float32), tr, C, T))
    plan, out = run_units(eng, logps, cands, means, seg0=seg0, mode=mode)
    assert eng.last_mode == ("fused" if mode == "auto" else mode)
    for u in range(plan.U):
        ref = oracle_unit(logps[u], cands[u][0], means[u], 30, 2000, seg0)
        bs = out["bs"][plan.blk_off[u]:plan.blk_off[u + 1]]
        assert np.array_equal(bs, ref["bs"]), f"block scores differ for unit {u}"
        check_unit(plan, out, u, ref, logps[u].shape[0])


@pytest.mark.parametrize("fs,max_len,C", [(7, 91, 12), (1, 20, 5), (13, 200, 7), (30, 2000, 100), (10, 1000, 33)])
def test_other_sampling_and_class_counts(eng, fs, max_len, C, mode):
    """C = 5, 7, 33 take the direct scan (row not a multiple of 16 B); max_len % fs == 0 gives -inf rows."""
    rng = np.random.default_rng(fs * 1000 + C)
    J = max_len // fs
    logps, cands, means = [], [], []
    for i in range(10):
        N = int(rng.integers(1, 7))
        K = int(rng.integers(max(N - 1, 1), N * J + 1))
        T = K * fs + int(rng.integers(0, fs))
        tr = rng.integers(0, C, N).tolist()
        logps.append(np.log(rng.dirichlet(np.ones(C), T)).astype([np.float32, np.float64][i % 2]))
        cands.append([tr])
        means.append(rng.uniform(0.6, max(T, 2), C))
    for dt in (np.float32, np.float64):
        idx = [i for i in range(10) if logps[i].dtype == dt]
        plan, out = run_units(eng, [logps[i] for i in idx], [cands[i] for i in idx], [means[i] for i in idx],
                              fs, max_len, seg0=(dt == np.float32), mode=mode)
        for u, i in enumerate(idx):
            ref = oracle_unit(logps[i], cands[i][0], means[i], fs, max_len, dt == np.float32)
            check_unit(plan, out, u, ref, logps[i].shape[0])


def test_rows_path_equals_params_path(eng):
    rng = np.random.default_rng(5)
    tr = [3, 1, 4, 1, 5]
    lp, _ = synth.planted_logp(rng, 1700, 9, tr, np.float32)
    m = rng.uniform(50, 600, 9)
    _, a = run_units(eng, [lp], [[tr]], [m], use_rows=False)
    _, b = run_units(eng, [lp], [[tr]], [m], use_rows=True)
    for k in ("score", "labels", "seg_blocks", "bp", "final_j"):
        assert np.array_equal(a[k], b[k])


def test_edge_cases_status(eng, mode):
    rng = np.random.default_rng(9)
    C = 6
    mk = lambda T: np.log(rng.dirichlet(np.ones(C), T)).astype(np.float32)
    logps = [mk(3990), mk(100), mk(3960), mk(59)]
    cands = [[[0, 1]], [[0, 1, 2, 3, 4, 5]], [[1, 2]], [[3]]]
    means = [np.full(C, 500.0)] * 4
    plan, out = run_units(eng, logps, cands, means, seg0=True, mode=mode)
    assert out["status"].tolist() == [1, 2, 0, 0]  # K > N*J infeasible; K < N short; K == N*J ok; K == 1 ok
    assert np.isnan(out["score"][0]) and out["score"][1] == -np.inf
    for u in (1, 2, 3):
        ref = oracle_unit(logps[u], cands[u][0], means[u], 30, 2000, True)
        check_unit(plan, out, u, ref, logps[u].shape[0])
    assert out["seg_blocks"][plan.tr_off[2]:plan.tr_off[3]].tolist() == [66, 66]


@pytest.mark.parametrize("dtype,seg0", [(np.float32, True), (np.float64, False)])
def test_generic_kernel_equals_register_kernel(eng, dtype, seg0):
    """The workspace kernel (J > 128 / N > 65) computes the same program: force it on shapes the
    register-resident kernels also cover and compare every output, edge cases included."""
    rng = np.random.default_rng(77)
    C = 12
    logps, cands, means = [], [], []
    for i in range(40):
        N = int(rng.integers(1, 9))
        T = int(rng.integers(40, 2600))
        tr = list(map(int, rng.integers(0, C, N)))
        lp, _ = synth.planted_logp(rng, T, C, tr, dtype)
        logps.append(lp)
        cands.append([tr])
        means.append(synth.class_means(rng.dirichlet(np.ones(N)).astype(np.float32), tr, C, T))
    mk = lambda T: np.log(rng.dirichlet(np.ones(C), T)).astype(dtype)
    logps += [mk(3990), mk(100), mk(3960), mk(59)]
    cands += [[[0, 1]], [[0, 1, 2, 3, 4, 5]], [[1, 2]], [[3]]]
    means += [np.full(C, 500.0)] * 4
    pa, a = run_units(eng, logps, cands, means, seg0=seg0, mode="split")
    pb, b = run_units(eng, logps, cands, means, seg0=seg0, force_generic=True)
    assert eng.last_mode == "generic"
    assert a["status"].tolist() == b["status"].tolist()
    assert np.array_equal(a["score"], b["score"], equal_nan=True)
    for k in ("final_j", "seg_blocks"):
        assert np.array_equal(a[k], b[k]), k
    for u in range(pa.U):  # labels / back-pointers of an infeasible unit are not written
        if a["status"][u] == 1:
            continue
        T = logps[u].shape[0]
        assert np.array_equal(a["labels"][pa.lab_off[u]:pa.lab_off[u] + T], b["labels"][pb.lab_off[u]:pb.lab_off[u] + T])
        assert np.array_equal(a["bp"][pa.bp_off[u]:pa.bp_off[u + 1]], b["bp"][pb.bp_off[u]:pb.bp_off[u + 1]])


@pytest.mark.parametrize("fs,max_len", [(1, 400), (3, 1000)])
def test_generic_kernel_large_J_vs_oracle(eng, fs, max_len):
    rng = np.random.default_rng(78)
    C = 8
    logps, cands, means = [], [], []
    for i in range(6):
        N = int(rng.integers(1, 7))
        T = int(rng.integers(5 * fs, 900 * fs // 2))
        tr = list(map(int, rng.integers(0, C, N)))
        lp, _ = synth.planted_logp(rng, T, C, tr, np.float32)
        logps.append(lp)
        cands.append([tr])
        means.append(rng.uniform(0.1, 0.6, C) * max_len)
    plan, out = run_units(eng, logps, cands, means, fs=fs, max_len=max_len, seg0=True)
    assert eng.last_mode == "generic"
    for u in range(len(logps)):
        T = logps[u].shape[0]
        if T // fs > len(cands[u][0]) * (max_len // fs):
            assert out["status"][u] == 1
            continue
        ref = oracle_unit(logps[u], cands[u][0], means[u], fs, max_len, True)
        check_unit(plan, out, u, ref, T)


def test_candidates_best_equals_argmax_of_singles(eng):
    rng = np.random.default_rng(13)
    logps, cands, means = [], [], []
    for v in range(6):
        base = rng.integers(0, 20, int(rng.integers(3, 9))).tolist()
        T = int(rng.integers(900, 4000))
        lp, _ = synth.planted_logp(rng, T, 20, base, np.float32)
        K = T // 30
        cl = synth.random_edits(rng, base, 20, 16, max(2, -(-K // 66)), min(30, K))
        logps.append(lp)
        cands.append(cl)
        means.append(synth.class_means(rng.dirichlet(np.ones(len(base))).astype(np.float32), base, 20, T))
    plan, out = run_units(eng, logps, cands, means, seg0=True, labels="best")
    assert plan.U == 96 and not plan.single
    for v in range(6):
        refs = [oracle_unit(logps[v], tr, means[v], 30, 2000, True) for tr in cands[v]]
        scores = np.array([r["score"] for r in refs])
        u0 = plan.cand_off[v]
        assert np.array_equal(out["score"][u0:u0 + 16], scores)
        b = int(np.argmax(scores))  # first maximum == lowest index wins
        assert out["best"][v] == u0 + b
        T = logps[v].shape[0]
        assert np.array_equal(out["labels"][plan.vid_off[v]:plan.vid_off[v] + T], refs[b]["labels"])


def test_long_video_stress_c4(eng, mode):
    """c4: T=40000, C=100, N=60 (K=1333, 3960 live states), float32 and float64."""
    rng = np.random.default_rng(4)
    tr = rng.permutation(100)[:60].tolist()
    lp, _ = synth.planted_logp(rng, 40000, 100, tr, np.float32)
    m = synth.class_means(rng.dirichlet(5 * np.ones(60)).astype(np.float32), tr, 100, 40000)
    for arr, seg0 in ((lp, True), (lp.astype(np.float64), False)):
        plan, out = run_units(eng, [arr], [[tr]], [m], seg0=seg0, mode=mode)
        check_unit(plan, out, 0, oracle_unit(arr, tr, m, 30, 2000, seg0), 40000)


def test_drop_in_viterbi_class(eng):
    """Same call sequence as the evaluator (reference src/mucon/evaluators.py:80,148,167,178-180)."""
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.viterbi import Viterbi
    g = load_golden("c1_f32")
    dec = Viterbi(None, None, frame_sampling=30, np_mode="numpy2")
    dec.grammar = SingleTranscriptGrammar(g["transcripts"][0], 48)
    dec.length_model = PoissonModel(g["means"])
    dec.set_multi_length(False)
    score, labels, segs = dec.decode(g["logp"])
    assert isinstance(labels, list) and len(labels) == 2000
    assert same_score(score, g["score"])
    assert labels == g["labels"].tolist()
    assert [(s.label, s.length) for s in segs] == g["segments"]
    # infeasible input: the reference raises AttributeError (K > N*J) / IndexError (T < fs)
    dec.grammar = SingleTranscriptGrammar([0, 1], 48)
    with pytest.raises(AttributeError):
        dec.decode(np.zeros((3990, 48), dtype=np.float32))
    with pytest.raises(IndexError):
        dec.decode(np.zeros((20, 48), dtype=np.float32))


def test_drop_in_viterbi_class_default_frame_sampling(eng):
    """Viterbi(grammar, length_model) with the reference's own default frame_sampling = 1
    (viterbi.py:34): J = max_length blocks, served by the generic kernel."""
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.viterbi import Viterbi
    g = load_golden("generic_fs1_J300")
    dec = Viterbi(SingleTranscriptGrammar(g["transcripts"][0], g["logp"].shape[1]),
                  PoissonModel(g["means"], max_length=int(g["max_len"])))
    score, labels, segs = dec.decode(g["logp"])
    assert same_score(score, g["score"])
    assert labels == g["labels"].tolist()
    assert [(s.label, s.length) for s in segs] == g["segments"]


def test_breakfast_split_properties_full_size(eng, mode):
    """c2 at full size (1712 videos): size-independent properties + oracle spot checks."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan
    T, trs = synth.breakfast_split(seed=0)
    rng = np.random.default_rng(1)
    C = 48
    V = len(T)
    total = int(T.sum())
    logp = torch.randn(total, C, device=eng.device, generator=torch.Generator(eng.device).manual_seed(3))
    logp = torch.log_softmax(logp, dim=1).contiguous()
    means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, C, int(t))
                      for tr, t in zip(trs, T)])
    plan = AlignPlan(T, [[tr.tolist()] for tr in trs], C, device=eng.device, len_params=poisson_params(means))
    eng.run(plan, logp, seg0_f32=True, mode=mode)
    torch.cuda.synchronize()
    assert eng.last_mode == ("fused" if mode == "auto" else mode)
    out = eng.fetch(plan, want_bp=False)
    assert (out["status"] == 0).all()
    K = T // 30
    sb_sum = np.add.reduceat(out["seg_blocks"], plan.tr_off[:-1])
    assert np.array_equal(sb_sum, K)  # segment lengths cover every block exactly once
    assert (out["seg_blocks"] >= 1).all() and (out["seg_blocks"] <= 66).all()
    # labels are the run-length expansion of (transcript, seg_blocks) with the leftover frames first
    for v in rng.choice(V, 40, replace=False):
        a, b = plan.tr_off[v], plan.tr_off[v + 1]
        rem = int(T[v] - 30 * K[v])
        exp = np.concatenate([np.full(rem, trs[v][-1])] + [np.full(30 * n, l) for l, n in zip(trs[v], out["seg_blocks"][a:b])])
        assert np.array_equal(out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]], exp)
    host = logp.cpu().numpy()
    longest = np.argsort(-T)[:6]  # these go through the wide launch (a warp per segment) in auto mode
    for v in list(rng.choice(V, 25, replace=False)) + list(longest):
        lp = host[plan.vid_off[v]:plan.vid_off[v + 1]]
        ref = oracle_unit(lp, trs[v].tolist(), means[v], 30, 2000, True)
        assert same_score(out["score"][v], ref["score"])
        assert np.array_equal(out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]], ref["labels"])
    if mode == "auto":
        # the long-tail split (two concurrent launches) changes nothing in the results
        assert plan.n_long > 0
        plan1 = AlignPlan(T, [[tr.tolist()] for tr in trs], C, device=eng.device, len_params=poisson_params(means),
                          long_K=0)
        assert plan1.n_long == 0
        eng.run(plan1, logp, seg0_f32=True, mode=mode)
        torch.cuda.synchronize()
        out1 = eng.fetch(plan1, want_bp=True)
        outb = eng.fetch(plan, want_bp=True)
        for k in ("score", "labels", "seg_blocks", "final_j", "status", "bp"):
            assert np.array_equal(outb[k], out1[k]), k


def test_empty_batch_and_single_frame_blocks(eng):
    """Degenerate batches: no videos at all; videos of exactly one block; a one-segment transcript."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan
    plan = AlignPlan([], [], 8, device=eng.device, len_params=np.zeros((0, 8, 3)))
    eng.run(plan, torch.zeros((0, 8), device=eng.device))
    torch.cuda.synchronize()
    out = eng.fetch(plan)
    assert out["score"].shape == (0,) and out["labels"].shape == (0,)
    rng = np.random.default_rng(4)
    logps = [np.log(rng.dirichlet(np.ones(8), t)).astype(np.float32) for t in (30, 59, 31, 1980)]
    cands = [[[3]], [[5]], [[0]], [[7]]]
    means = [np.full(8, 40.0)] * 4
    for m in ("auto", "split", "lanes"):
        plan, out = run_units(eng, logps, cands, means, seg0=True, mode=m)
        assert out["status"].tolist() == [0, 0, 0, 0]
        for u in range(4):
            ref = oracle_unit(logps[u], cands[u][0], means[u], 30, 2000, True)
            check_unit(plan, out, u, ref, logps[u].shape[0])


@pytest.mark.parametrize("dtype,C", [(np.float32, 48), (np.float64, 48), (np.float32, 20), (np.float32, 100)])
def test_pooled_source_is_bit_identical_to_expanded(eng, dtype, C):
    """mucon_viterbi_align_fused_pooled: the scan reads the [Tz, C] table at the backbone's pooled resolution through
    the nearest-neighbour index of F.interpolate instead of the expanded [T, C] array.  Block scores, scores,
    back-pointers and labels must equal, bit for bit, the run on the expanded array (and the oracle on it)."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan
    rng = np.random.default_rng(7)
    tabs, logps, cands, means, Tzs = [], [], [], [], []
    for i in range(40):
        N = int(rng.integers(1, 13))
        K = int(rng.integers(max(N, 1), min(N * 66, 340) + 1))
        T = K * 30 + int(rng.integers(0, 30))
        Tz = max(1, T // 16) if i % 4 else max(1, int(rng.integers(1, T + 1)))   # also odd ratios, incl. Tz == T
        tr = rng.integers(0, C, N).tolist()
        lp, _ = synth.planted_logp(rng, Tz, C, tr, dtype)
        idx = np.minimum(np.floor(np.arange(T, dtype=np.float32) * (np.float32(Tz) / np.float32(T))).astype(np.int64),
                         Tz - 1)
        tabs.append(lp)
        logps.append(np.ascontiguousarray(lp[idx]))
        Tzs.append(Tz)
        cands.append([tr])
        means.append(synth.class_means(rng.dirichlet(np.ones(N)).astype(np.float32), tr, C, T))
    seg0 = dtype == np.float32
    params = np.stack([poisson_params(m) for m in means])
    T_all = [l.shape[0] for l in logps]
    outs = []
    for pooled in (False, True):
        plan = AlignPlan(T_all, cands, C, device=eng.device, labels="all", len_params=params)
        if pooled:
            src = torch.from_numpy(np.concatenate(tabs)).to(eng.device)
            z_off = torch.from_numpy(np.concatenate([[0], np.cumsum(Tzs)]).astype(np.int64)).to(eng.device)
            eng.run(plan, src, seg0_f32=seg0, z_off=z_off)
        else:
            eng.run(plan, torch.from_numpy(np.concatenate(logps)).to(eng.device), seg0_f32=seg0, mode="fused")
        torch.cuda.synchronize()
        assert eng.last_mode == "fused"
        outs.append((plan, eng.fetch(plan, want_bp=True)))
    (pa, a), (pb, b) = outs
    for k in ("score", "final_j", "status", "seg_blocks", "labels", "bp", "bs"):
        assert np.array_equal(a[k], b[k]), k
    for u in range(0, pa.U, 5):
        ref = oracle_unit(logps[u], cands[u][0], means[u], 30, 2000, seg0)
        check_unit(pb, b, u, ref, logps[u].shape[0])


def test_c3_full_size_candidate_sets(eng):
    """config c3 at FULL size: the 1712-video split x 64 candidate transcripts (109 568 units) from array-form
    candidates (FlatCandidates): per video the winner is the first maximum of its candidates' scores, its labels are
    the run-length expansion of the winner's segment lengths, and every candidate of six sampled videos (the longest
    two among them) equals the oracle bit for bit."""
    import bench
    import bench_legs
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, FlatCandidates
    T, trs, _ = bench.make_split(0)
    V = len(T)
    cands, means = bench_legs.c3_candidates(T, trs, range(V))
    logp = bench.device_logp(T, trs, 0, eng.device)
    plan = AlignPlan(T, FlatCandidates.from_lists(cands), 48, device=eng.device, len_params=poisson_params(means),
                     labels="best")
    assert plan.U == 64 * V
    eng.run(plan, logp, seg0_f32=True)
    torch.cuda.synchronize()
    assert eng.last_mode == "lanes"
    out = eng.fetch(plan)
    assert np.isin(out["status"], (0, 2)).all()
    sc = out["score"].reshape(V, 64)
    # the maximum wins; EXACT ties (this synthetic split has them: candidates that differ in a last one-block segment
    # whose two classes share a log-probability) go to the hypothesis latest in the reference's dict order
    # (mucon_viterbi_select_ranked; tests/test_ties.py holds the reference-minted cases), duplicates to the lowest index
    first = np.arange(V) * 64 + np.argmax(sc, axis=1)
    ld = (out["final_j"] + np.diff(plan.tr_off)).reshape(V, 64)
    r0, r1 = plan.tie_rank[:, 0].reshape(V, 64), plan.tie_rank[:, 1].reshape(V, 64)
    want = np.empty(V, dtype=np.int64)
    for v in range(V):
        top = np.nonzero(sc[v] == sc[v].max())[0]
        want[v] = 64 * v + max(top, key=lambda c: (r0[v, c], ld[v, c], r1[v, c], -c))
    assert np.array_equal(out["best"], want)
    moved = np.nonzero(want != first)[0]
    for v in moved[np.argsort(T[moved])][:2]:      # ... and that IS what the reference's table does (oracle, all 64 at once)
        lp = logp[plan.vid_off[v]:plan.vid_off[v + 1]].cpu().numpy()
        s, labels, _ = hyp_viterbi.decode(lp, cands[v], poisson.poisson_table(means[v], 2000), 2000, 30)
        assert same_score(out["score"][want[v]], s)
        assert np.array_equal(out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]], np.asarray(labels, dtype=np.int32)), v
    K = T // 30
    rng = np.random.default_rng(2)
    sample = list(rng.choice(V, 4, replace=False)) + list(np.argsort(-T)[:2])
    for v in sample:
        u = int(out["best"][v])
        a, b = plan.tr_off[u], plan.tr_off[u + 1]
        tr = cands[v][u - 64 * v]
        rem = int(T[v] - 30 * K[v])
        exp = np.concatenate([np.full(rem, tr[-1])] + [np.full(30 * n, l) for l, n in zip(tr, out["seg_blocks"][a:b])])
        assert np.array_equal(out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]], exp)
    host = {int(v): logp[plan.vid_off[v]:plan.vid_off[v + 1]].cpu().numpy() for v in sample}
    for v in sample:
        for c in range(64):
            u = 64 * int(v) + c
            rows = coracle.poisson_rows(poisson.poisson_params(means[v])[cands[v][c]], 30, 2000)
            ref = coracle.viterbi(coracle.block_scores(host[int(v)], 30), cands[v][c], rows, True)
            assert same_score(out["score"][u], ref["score"]), (v, c)
            assert np.array_equal(out["seg_blocks"][plan.tr_off[u]:plan.tr_off[u + 1]], ref["seg_blocks"]), (v, c)


@pytest.mark.parametrize("dist_name", ["equal", "bimodal", "few_long"])
def test_other_length_distributions(eng, dist_name):
    """The long-tail launch policy (the videos within 15 % of the longest go to a wide launch of their own) was tuned on
    the log-normal c2 split: on equal-length, bimodal and few-long-video batches the automatic policy, the forced
    single launch (long_K = 0) and a forced split must give identical results, equal to the oracle."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan
    rng = np.random.default_rng(17)
    V, C = 400, 48
    if dist_name == "equal":
        T = np.full(V, 2250)
    elif dist_name == "bimodal":
        T = np.where(rng.random(V) < 0.5, rng.integers(500, 800, V), rng.integers(8500, 9500, V))
    else:
        T = np.concatenate([rng.integers(300, 2000, V - 5), rng.integers(9000, 10000, 5)])
    trs = []
    for t in T:
        K = int(t) // 30
        n = int(rng.integers(max(2, -(-K // 66)), min(12, K) + 1))
        trs.append(rng.integers(0, C, n))
    means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, C, int(t))
                      for tr, t in zip(trs, T)])
    logp = torch.log_softmax(torch.randn(int(T.sum()), C, device=eng.device,
                                         generator=torch.Generator(eng.device).manual_seed(5)), dim=1).contiguous()
    outs = {}
    for name, long_K in (("auto", None), ("single", 0), ("split_half", int(0.5 * (T // 30).max()))):
        plan = AlignPlan(T, [[tr.tolist()] for tr in trs], C, device=eng.device, len_params=poisson_params(means),
                         long_K=long_K)
        eng.run(plan, logp, seg0_f32=True, write_bs=False)
        torch.cuda.synchronize()
        assert eng.last_mode == "fused"
        outs[name] = (plan, eng.fetch(plan))
    ref_plan, ref = outs["single"]
    assert (ref["status"] == 0).all()
    for name in ("auto", "split_half"):
        for k in ("score", "seg_blocks", "labels", "final_j"):
            assert np.array_equal(outs[name][1][k], ref[k]), (name, k)
    host = logp.cpu().numpy()
    for v in list(rng.choice(V, 10, replace=False)) + [int(np.argmax(T))]:
        lp = host[ref_plan.vid_off[v]:ref_plan.vid_off[v + 1]]
        o = oracle_unit(lp, trs[v].tolist(), means[v], 30, 2000, True)
        assert same_score(ref["score"][v], o["score"])
        assert np.array_equal(ref["labels"][ref_plan.vid_off[v]:ref_plan.vid_off[v + 1]], o["labels"])


@pytest.mark.gpu
def test_peer_exchange_two_gpus():
    """dist.PeerExchange: result stores repeated into the peers' receive buffers from the kernel epilogue equal an NCCL
    all_gather of the payloads (needs two GPUs on the node; skipped otherwise)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(root, "scripts", "peer_exchange_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and '"peer_exchange_ok": true' in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_drop_in_decode_with_float32_mean_lengths(eng):
    """a PoissonModel built from float32 means (rows in the reference's dtypes, not the float64 parameter triple):
    score / labels / segments equal the reference's frozen decode"""
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.viterbi import Viterbi
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poisson_f32.npz"))
    if str(g["numpy_version"]).split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("minted under another NumPy major version")
    dec = Viterbi(SingleTranscriptGrammar(g["transcript"].tolist(), 48), PoissonModel(g["means"]), frame_sampling=30,
                  device=eng.device)
    score, labels, segs = dec.decode(g["logp"])
    assert same_score(score, g["score"])
    assert labels == g["labels"].tolist()
    assert [(s.label, s.length) for s in segs] == [tuple(x) for x in g["segs"].tolist()]


def test_host_align_pipeline_equals_serial_runs(eng):
    """viterbi.HostAlignPipeline (H2D of batch i+1 under the kernels and the D2H of batch i, three streams, two input
    buffers): every batch's labels / scores / segment lengths equal a serial run of the same batch, also when the batch
    shape changes from one submit to the next and when more batches than slots are in flight over time."""
    from mucon_b200 import _lib
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, HostAlignPipeline
    rng = np.random.default_rng(23)
    C = 48
    batches = []
    for b in range(5):
        V = 40 if b != 3 else 17
        T = rng.integers(300, 4000, V)
        trs = [rng.integers(0, C, int(rng.integers(2, 9))).tolist() for _ in range(V)]
        means = np.stack([synth.class_means(rng.dirichlet(np.ones(len(tr))).astype(np.float32), tr, C, int(t))
                          for tr, t in zip(trs, T)])
        logp = torch.log_softmax(torch.randn(int(T.sum()), C, generator=torch.Generator().manual_seed(b)), 1).contiguous()
        batches.append((T, trs, means, logp.pin_memory()))

    def mk(T, trs, means):
        return lambda: AlignPlan(T, [[tr] for tr in trs], C, device=eng.device, len_params=poisson_params(means))

    want = []
    for T, trs, means, logp in batches:
        plan = mk(T, trs, means)()
        eng.run(plan, logp.to(eng.device), seg0_f32=True)
        out = eng.fetch(plan)
        want.append((out["labels"].copy(), out["score"].copy(), out["seg_blocks"].copy()))
    pipe = HostAlignPipeline(eng)
    prev = None
    got = {}
    for b, (T, trs, means, logp) in enumerate(batches):
        tk = pipe.submit(logp, mk(T, trs, means), seg0_f32=True)
        if prev is not None:
            _, la, sc, sg = pipe.result(prev)
            got[prev] = (la.numpy().copy(), sc.numpy().copy(), sg.numpy().copy())
        prev = tk
    _, la, sc, sg = pipe.result(prev)
    got[prev] = (la.numpy().copy(), sc.numpy().copy(), sg.numpy().copy())
    for b in range(len(batches)):
        for a, w in zip(got[b], want[b]):
            assert np.array_equal(a, w), b
    # a slot cannot be reused before its results were collected
    t0 = pipe.submit(batches[0][3], mk(*batches[0][:3]), seg0_f32=True)
    t1 = pipe.submit(batches[1][3], mk(*batches[1][:3]), seg0_f32=True)
    with pytest.raises(_lib.MuconError):
        pipe.submit(batches[2][3], mk(*batches[2][:3]), seg0_f32=True)
    pipe.result(t0)
    pipe.result(t1)
