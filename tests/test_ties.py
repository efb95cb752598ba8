"""Exact score ties between candidate transcripts: the reference returns the final hypothesis that is LATEST in its
insertion-ordered hypothesis dict (viterbi.py:26-28,93-138).  tests/golden/ties.npz holds 160 tied cases decoded by the
unmodified reference (make_golden_ties.py; "lowest index wins" is wrong in 72 of them)."""
import os

import numpy as np
import pytest

from oracle import hyp_viterbi, poisson

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    G = np.load(os.path.join(HERE, "golden", "ties.npz"))
    for i in range(int(G["n_cases"])):
        k = f"t{i}_"
        fs, max_len, C = (int(x) for x in G[k + "fs_maxlen_C"])
        cuts = np.cumsum(G[k + "tr_len"])[:-1]
        trs = [t.tolist() for t in np.split(G[k + "tr"], cuts)]
        yield dict(i=i, logp=G[k + "logp"], means=G[k + "means"], fs=fs, max_len=max_len, C=C, trs=trs,
                   score=float(G[k + "score"]), labels=G[k + "labels"],
                   segments=list(zip(G[k + "seg_label"].tolist(), G[k + "seg_length"].tolist())))


def test_oracle_follows_the_reference_dict_order():
    """pins oracle/hyp_viterbi.py (table order, successor-set iteration order) to the reference on tied inputs"""
    for c in cases():
        tab = poisson.poisson_table(c["means"], c["max_len"])
        s, labels, segs = hyp_viterbi.decode(c["logp"], c["trs"], tab, c["max_len"], c["fs"])
        assert s == c["score"], c["i"]
        assert np.array_equal(np.asarray(labels, dtype=np.int32), c["labels"]), c["i"]
        assert segs == c["segments"], c["i"]


def test_tie_rank_rule_predicts_the_reference_winner():
    """host logic without a GPU: single-transcript decodes (oracle) + grammar.tie_ranks + the device's key order
    (score, rank of the first label, last segment blocks + transcript length, rank of the rest) = the reference"""
    from mucon_b200.grammar import ModifiedPathGrammar, tie_ranks, tie_ranks_for_lists
    for c in cases():
        fs, T = c["fs"], c["logp"].shape[0]
        tab = poisson.poisson_table(c["means"], c["max_len"])
        g = ModifiedPathGrammar(c["trs"], c["C"])
        ranks = tie_ranks(g.successors, c["trs"], g.start_symbol())
        assert np.array_equal(ranks, tie_ranks_for_lists(c["trs"]))
        best = None
        for u, tr in enumerate(c["trs"]):
            try:
                s, labels, segs = hyp_viterbi.decode(c["logp"], [tr], tab, c["max_len"], fs)
            except Exception:
                continue
            if not np.isfinite(s):
                continue
            last_blocks = (segs[-1][1] - (T - (T // fs) * fs)) // fs
            key = (s, int(ranks[u, 0]), last_blocks + len(tr), int(ranks[u, 1]), -u)
            if best is None or key > best[0]:
                best = (key, labels)
        assert best[0][0] == c["score"], c["i"]
        assert np.array_equal(np.asarray(best[1], dtype=np.int32), c["labels"]), c["i"]


@pytest.mark.gpu
def test_drop_in_decode_returns_the_reference_candidate_on_ties():
    from mucon_b200 import ModifiedPathGrammar, PoissonModel
    from mucon_b200.viterbi import Viterbi
    for c in cases():
        dec = Viterbi(ModifiedPathGrammar(c["trs"], c["C"]), PoissonModel(c["means"], max_length=c["max_len"]),
                      frame_sampling=c["fs"])
        score, labels, segs = dec.decode(c["logp"])
        assert float(score) == c["score"], c["i"]
        assert labels == c["labels"].tolist(), c["i"]
        assert [(s.label, s.length) for s in segs] == c["segments"], c["i"]


@pytest.mark.gpu
@pytest.mark.parametrize("flat", [False, True])
def test_batched_plan_returns_the_reference_candidate_on_ties(flat):
    """all cases of one (fs, max_len, C) shape in ONE plan: candidate arg-max per video on the device"""
    import torch
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import AlignPlan, FlatCandidates, ViterbiEngine
    eng = ViterbiEngine("cuda:0")
    groups = {}
    for c in cases():
        groups.setdefault((c["fs"], c["max_len"], c["C"]), []).append(c)
    for (fs, max_len, C), cs in groups.items():
        T = np.array([c["logp"].shape[0] for c in cs], dtype=np.int64)
        cands = [c["trs"] for c in cs]
        plan = AlignPlan(T, FlatCandidates.from_lists(cands) if flat else cands, C, fs=fs, max_len=max_len,
                         device=eng.device, labels="best", len_params=poisson_params(np.stack([c["means"] for c in cs])))
        logp = torch.from_numpy(np.concatenate([c["logp"] for c in cs])).to(eng.device)
        eng.run(plan, logp, seg0_f32=False)
        out = eng.fetch(plan)
        off = np.concatenate([[0], np.cumsum(T)])
        for v, c in enumerate(cs):
            u = int(out["best"][v])
            assert float(out["score"][u]) == c["score"], c["i"]
            assert np.array_equal(out["labels"][off[v]:off[v + 1]], c["labels"]), c["i"]
