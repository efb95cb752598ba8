"""mucon_b200 -- B200-native (sm_100a) implementation of MuCon's data-parallel hot path.

Host-side mirror of the reference interfaces for that path only:
    viterbi.Viterbi, grammar.*Grammar, length_model.PoissonModel   (reference src/core/viterbi/*)
    masks.create_masks / project_lengths_softmax                   (reference src/mucon/masks.py)
    temporal.WaveNetBlock + model.* forward helpers                (reference src/core/modules/temporal.py,
                                                                    src/mucon/models.py:360-374,567-582,746-773)
    train.forward_train_packed / TrainStep, loss.mucon_loss*       (trainers.py:125-131, models.py:398-525)
    shead.SHead, inference.infer_and_align                         (models.py:585-745; evaluators.py:128-180)
    evaluate.align_videos, metrics.*                               (evaluators.py:147-243)
All compute lives in libmucon_b200.so (C ABI: include/mucon_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .grammar import Grammar, ModifiedPathGrammar, PathGrammar, SingleTranscriptGrammar  # noqa: F401
from .length_model import LengthModel, PoissonModel  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import mucon_b200` stays cheap
    if name in ("Viterbi", "ViterbiEngine", "AlignPlan"):
        from . import viterbi
        return getattr(viterbi, name)
    raise AttributeError(name)
