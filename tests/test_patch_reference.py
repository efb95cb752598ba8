"""patch_reference() against the real reference tree (only where /root/reference exists -- this container; the GPU
box does not have it, there the numeric side of the same flow is covered by tests/golden/evaluator_flow.npz, minted
from the unmodified reference evaluator).  Checks the wiring: after the patch the reference's own import paths hand
out this package's classes, the reference evaluator module sees them, a reference-constructed WaveNetBlock converts
with its weights, and -- this container has no GPU -- the patched entry points fail loudly instead of falling back."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")


@pytest.fixture()
def ref_modules():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden_loss  # noqa: F401  installs the fandak / yacs stubs and imports the reference's mucon.models
    import core.modules.temporal as rtemporal
    import core.viterbi.grammar as rgrammar
    import core.viterbi.length_model as rlm
    import core.viterbi.viterbi as rviterbi
    import mucon.masks as rmasks
    from mucon_b200 import patch
    yield dict(temporal=rtemporal, grammar=rgrammar, lm=rlm, viterbi=rviterbi, masks=rmasks, patch=patch)
    patch.unpatch_reference()
    sys.path.remove(here)


def test_patch_installs_and_restores(ref_modules):
    import mucon_b200
    from mucon_b200 import temporal, viterbi
    m = ref_modules
    orig = (m["viterbi"].Viterbi, m["masks"].create_masks, m["temporal"].WaveNetBlock)
    done = m["patch"].patch_reference()
    for pair in (("core.viterbi.viterbi", "Viterbi"), ("core.viterbi.length_model", "PoissonModel"),
                 ("core.viterbi.grammar", "SingleTranscriptGrammar"), ("mucon.masks", "create_masks"),
                 ("mucon.models", "create_masks"), ("core.modules.temporal", "WaveNetBlock"),
                 ("mucon.models", "WaveNetBlock")):
        assert pair in done, pair
    assert m["viterbi"].Viterbi is viterbi.Viterbi
    assert m["temporal"].WaveNetBlock is temporal.WaveNetBlock
    import mucon.models as rmodels
    assert rmodels.create_masks is mucon_b200.masks.create_masks and rmodels.WaveNetBlock is temporal.WaveNetBlock
    m["patch"].unpatch_reference()
    assert (m["viterbi"].Viterbi, m["masks"].create_masks, m["temporal"].WaveNetBlock) == orig


def test_patched_entry_points_route_to_the_gpu_path(ref_modules):
    """the call sequence of evaluators.py:80,148,167,178 through the REFERENCE's module paths after the patch"""
    from mucon_b200 import _lib
    m = ref_modules
    m["patch"].patch_reference()
    dec = m["viterbi"].Viterbi(None, None, frame_sampling=30)
    dec.grammar = m["grammar"].SingleTranscriptGrammar([0, 5, 7, 5, 12, 0], 48)
    dec.length_model = m["lm"].PoissonModel(np.full(48, 300.0))
    dec.set_multi_length(False)
    logp = np.log(np.full((2000, 48), 1.0 / 48, dtype=np.float32))
    if torch.cuda.is_available():
        score, labels, segs = dec.decode(logp)
        assert len(labels) == 2000 and sum(s.length for s in segs) == 2000
    else:
        with pytest.raises((_lib.MuconError, RuntimeError, AssertionError)):
            dec.decode(logp)   # no CUDA device: the product path must fail, not fall back
        with pytest.raises(_lib.MuconError):
            m["masks"].create_masks(T=100, L=torch.tensor([40.0, 60.0]), template="box", overlap=0.0)


def test_patch_model_converts_a_reference_block(ref_modules):
    """a reference-constructed WaveNetBlock -> this package's block, same weights (state_dict round trip)"""
    m = ref_modules

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ft = m["temporal"].WaveNetBlock(in_channels=64, out_dims=128, stages=[1, 2, 4], pooling_layers=[1])

    from mucon_b200 import temporal
    h = Holder().eval()
    ref_sd = {k: v.clone() for k, v in h.ft.state_dict().items()}
    m["patch"].patch_model(h)
    assert isinstance(h.ft, temporal.WaveNetBlock) and not h.ft.training
    assert h.ft.stages == [1, 2, 4] and h.ft.pooling_layers == [1]
    for k, v in h.ft.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
