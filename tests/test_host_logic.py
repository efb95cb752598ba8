"""Host-side logic that needs no GPU: array-form candidate sets, the c3 sharding used by bench.py, the train plan's
nearest-neighbour index table, the batched loss's index bookkeeping."""
import numpy as np
import pytest
import torch


def test_flat_candidates_equal_nested_lists():
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, FlatCandidates
    rng = np.random.default_rng(0)
    T = rng.integers(300, 4000, 23)
    cands = [[rng.integers(0, 48, int(rng.integers(2, 9))).tolist() for _ in range(int(rng.integers(1, 6)))] for _ in T]
    means = rng.uniform(30, 400, (len(T), 48))
    a = AlignPlan(T, cands, 48, device="cpu", len_params=poisson_params(means), labels="best")
    b = AlignPlan(T, FlatCandidates.from_lists(cands), 48, device="cpu", len_params=poisson_params(means), labels="best")
    for name in ("tr", "tr_off", "cand_off", "unit_vid", "warp_unit", "order_v", "bp_off", "lab_off"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.U == b.U and a.max_N == b.max_N and a.n_cta == b.n_cta and a.n_lane_warps == b.n_lane_warps
    with pytest.raises(ValueError):
        AlignPlan(T, FlatCandidates([1] * len(T), [3] * (len(T) - 1), [0] * 3 * (len(T) - 1)), 48, device="cpu",
                  len_params=poisson_params(means))


def test_c3_sharding_keeps_all_candidates_of_a_video_together():
    import bench
    from mucon_b200 import dist as mdist
    T, trs, _ = bench.make_split(0)
    for world in (2, 4, 8):
        shards = mdist.shard_videos(T, [64] * len(T), world)
        allv = np.sort(np.concatenate(shards))
        assert np.array_equal(allv, np.arange(len(T)))
        load = np.array([T[s].sum() for s in shards], dtype=np.float64)
        assert load.max() / load.mean() < 1.01      # greedy longest-first: within 1 % of perfect balance


def test_train_plan_expand_index_is_torch_nearest():
    """the frame -> pooled-row table of the training tail equals F.interpolate(mode='nearest')"""
    import torch.nn.functional as F
    from mucon_b200.temporal import BackbonePlan
    from mucon_b200.train import _plan_train_tables
    Ts = [700, 333, 17, 2048, 1999]
    plan = BackbonePlan(Ts, 4, "cpu")
    vid, idx, counts = _plan_train_tables(plan, "cpu")
    pos = 0
    for v, T in enumerate(Ts):
        Tz = int(plan.T[-1][v])
        want = F.interpolate(torch.arange(Tz, dtype=torch.float32)[None, None], T)[0, 0].long() + int(plan.off_host[-1][v])
        assert torch.equal(idx[pos:pos + T], want), v
        pos += T
    assert counts.tolist() == [float(t) for t in plan.T[-1]]


def test_smoothing_loss_packed_matches_reference_statements():
    import torch.nn.functional as F
    from mucon_b200.loss import smoothing_loss_packed
    g = torch.Generator().manual_seed(0)
    Ts = [40, 7, 130]
    segs = [torch.randn(t, 12, generator=g) * 3 for t in Ts]
    want = sum(torch.clamp(F.mse_loss(F.log_softmax(s, 1)[1:], F.log_softmax(s, 1)[:-1]), min=0.0, max=16.0) for s in segs) / 3
    got = smoothing_loss_packed(torch.cat(segs), Ts)
    assert abs(got.item() - want.item()) <= 1e-6 * max(1.0, abs(want.item()))


def test_both_bench_arms_see_the_same_arrays():
    """the reference arm's sample is a prefix of the arrays the GPU arm aligns"""
    import bench
    T, trs, _ = bench.make_split(0)
    jobs = bench.cpu_sample_jobs(3, seed=0)
    ours = bench.make_host_logp(T[:3], trs[:3], 0)
    pos = 0
    for v in range(3):
        assert np.array_equal(ours[pos:pos + int(T[v])], jobs[v][0])
        assert np.array_equal(jobs[v][1], trs[v])
        pos += int(T[v])


def test_reference_arm_json_contract():
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads (CPU only, bounded sample)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == bench_metric() and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def bench_metric():
    import bench
    return bench.METRIC


def test_both_arms_print_the_same_config():
    import bench
    assert bench.bench_config(1, "peer") == bench.bench_config(1, "peer")
    c = bench.bench_config(8, "peer")
    assert c["workload"] == bench.WORKLOAD and c["videos_per_gpu"] == 1712 and "larger than the 126 MB L2" in c["l2"]


def test_grammar_prefix_tree_is_lazy_and_iterates_like_the_reference():
    """The grammars build `successors` on first use (the evaluator constructs one per video and the CUDA path reads
    `candidates` only); when built, the sets are the reference's ({x}.union(old), grammar.py:150-154,185-189,203-207) --
    same members AND same iteration order, which decides exact ties between candidates (tests/test_ties.py)."""
    import os
    import random
    import sys
    from mucon_b200.grammar import ModifiedPathGrammar, SingleTranscriptGrammar
    g = SingleTranscriptGrammar([3, 7, 3], 10)
    assert g._succ is None and g.candidates == [[3, 7, 3]]
    assert g.possible_successors((-1, 3)) == {7} and g.score((-1,), 3) == 0.0 and g.score((-1,), 4) == -np.inf
    assert g._succ is not None
    g.successors = {(-1,): {1}}          # the attribute stays assignable, as on the reference's objects
    assert g.possible_successors((-1,)) == {1}
    if not os.path.isdir("/root/reference/src"):
        return
    sys.path.insert(0, "/root/reference/src")
    try:
        from core.viterbi.grammar import ModifiedPathGrammar as RefGrammar
    finally:
        sys.path.pop(0)
    rng = random.Random(5)
    for _ in range(300):
        C = rng.choice([5, 12, 48, 300])
        trs = [[rng.randrange(C) for _ in range(rng.randrange(1, 6))] for _ in range(rng.randrange(1, 10))]
        ours, ref = ModifiedPathGrammar(trs, C), RefGrammar([list(t) for t in trs], C)
        assert set(ours.successors) == set(ref.successors)
        for ctx, nxt in ref.successors.items():
            assert list(ours.successors[ctx]) == list(nxt), (trs, ctx)
