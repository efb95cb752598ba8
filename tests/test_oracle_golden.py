"""CPU: every oracle restatement reproduces the reference's frozen outputs (tests/golden)."""
import numpy as np
import pytest

from oracle import coracle, dense_viterbi, hyp_viterbi, poisson
from tests.util import golden_names, load_golden, same_score


@pytest.mark.parametrize("name", golden_names())
def test_hypothesis_port_matches_golden(name):
    g = load_golden(name)
    if g["seg0_f32"] != dense_viterbi.numpy_seg0_f32(g["logp"].dtype):
        pytest.skip("fixture minted under a different NumPy promotion regime")
    tab = poisson.poisson_table(g["means"], g["max_len"])
    s, labels, segs = hyp_viterbi.decode(g["logp"], g["transcripts"], tab, g["max_len"], g["fs"])
    assert same_score(s, g["score"])
    assert np.array_equal(np.asarray(labels, dtype=np.int32), g["labels"])
    assert segs == g["segments"]


def _best_single(g, decode_one):
    best = None
    for i, tr in enumerate(g["transcripts"]):
        try:
            d = decode_one(tr)
        except (dense_viterbi.Infeasible, ValueError):
            continue
        if best is None or d["score"] > best[1]["score"]:
            best = (i, d)
    return best


@pytest.mark.parametrize("name", golden_names())
def test_dense_restatement_matches_golden(name):
    g = load_golden(name)
    tab = poisson.poisson_table(g["means"], g["max_len"])

    def one(tr):
        rows = dense_viterbi.length_rows(tab, tr, g["fs"], g["max_len"])
        return dense_viterbi.decode(g["logp"], tr, rows, g["fs"], seg0_f32=g["seg0_f32"])

    i, d = _best_single(g, one)
    assert same_score(d["score"], g["score"])
    assert np.array_equal(d["labels"], g["labels"])
    tr = g["transcripts"][i]
    assert dense_viterbi.segments_from_blocks(d["seg_blocks"], tr, g["fs"], g["logp"].shape[0]) == g["segments"]
    if len(g["transcripts"]) == 1 and np.isfinite(g["score"]):
        assert np.array_equal(d["bp"], g["bp"])  # back-pointers, bit-exact


@pytest.mark.parametrize("name", golden_names())
def test_c_port_matches_golden(name):
    g = load_golden(name)
    params = poisson.poisson_params(g["means"])

    def one(tr):
        rows = coracle.poisson_rows(params[tr], g["fs"], g["max_len"])
        d = coracle.decode_video(g["logp"], tr, rows, g["fs"], g["seg0_f32"])
        bs = coracle.block_scores(g["logp"], g["fs"])
        d.update(coracle.viterbi(bs, tr, rows, g["seg0_f32"]))
        return d

    i, d = _best_single(g, one)
    assert same_score(d["score"], g["score"])
    assert np.array_equal(d["labels"], g["labels"])
    if len(g["transcripts"]) == 1 and np.isfinite(g["score"]):
        assert np.array_equal(d["bp"], g["bp"])


def test_poisson_rows_c_equals_numpy_table():
    rng = np.random.default_rng(3)
    means = rng.uniform(0.6, 5000, 40)
    for fs, max_len in [(30, 2000), (7, 91), (1, 20), (10, 2000)]:
        tab = poisson.poisson_table(means, max_len)
        tr = rng.integers(0, 40, 9).tolist()
        assert np.array_equal(dense_viterbi.length_rows(tab, tr, fs, max_len),
                              coracle.poisson_rows(poisson.poisson_params(means)[tr], fs, max_len))


def test_infeasible_inputs_raise():
    logp = np.zeros((3990, 3), dtype=np.float32)
    tab = poisson.poisson_table(np.full(3, 100.0))
    rows = dense_viterbi.length_rows(tab, [0, 1], 30, 2000)
    with pytest.raises(dense_viterbi.Infeasible):  # K=133 > 66*2 (SURVEY V-edge)
        dense_viterbi.decode(logp, [0, 1], rows)
    with pytest.raises(Exception):
        hyp_viterbi.decode(logp, [[0, 1]], tab)
    with pytest.raises(dense_viterbi.Infeasible):  # T < fs
        dense_viterbi.decode(logp[:20], [0, 1], rows)


def test_poisson_model_float32_means_follow_the_reference_dtypes():
    """PoissonModel built from float32 mean lengths: the reference's norms and l*log(m) - m run in float32
    (length_model.py:54-71); the table and the decoder rows must equal the frozen reference table bit for bit
    (tests/golden/poisson_f32.npz; minted under the NumPy version recorded in the file)."""
    import os
    import numpy as np
    import pytest
    from mucon_b200.length_model import PoissonModel
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poisson_f32.npz"))
    if str(g["numpy_version"]).split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("minted under another NumPy major version (different promotion rules)")
    lm = PoissonModel(g["means"])
    assert not lm.exact_params
    assert np.array_equal(lm.poisson[:40], g["table_head"], equal_nan=True)
    assert np.array_equal(lm.poisson[np.arange(1, 67) * 30], g["table_rows"], equal_nan=True)
    rows = lm.rows_for(g["transcript"], 30, 66)
    assert np.array_equal(rows, g["table_rows"][:, g["transcript"]].T, equal_nan=True)
    assert PoissonModel(g["means"].astype(np.float64)).exact_params
