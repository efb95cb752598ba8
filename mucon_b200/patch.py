"""patch_reference(): installs the GPU implementations into an importable reference tree
(`core.viterbi.viterbi.Viterbi`, `mucon.masks.create_masks`, ...) without editing its sources."""
import importlib


def patch_reference(viterbi=True, masks=True):
    """Returns the list of (module, attribute) pairs that were replaced."""
    done = []
    if viterbi:
        from . import grammar, length_model
        from .viterbi import Viterbi
        for modname, attr, obj in (
            ("core.viterbi.viterbi", "Viterbi", Viterbi),
            ("core.viterbi.length_model", "PoissonModel", length_model.PoissonModel),
            ("core.viterbi.grammar", "SingleTranscriptGrammar", grammar.SingleTranscriptGrammar),
            ("core.viterbi.grammar", "ModifiedPathGrammar", grammar.ModifiedPathGrammar),
            ("mucon.evaluators", "Viterbi", Viterbi),
            ("mucon.evaluators", "PoissonModel", length_model.PoissonModel),
            ("mucon.evaluators", "SingleTranscriptGrammar", grammar.SingleTranscriptGrammar),
        ):
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            if hasattr(mod, attr):
                setattr(mod, attr, obj)
                done.append((modname, attr))
    if masks:
        from .masks import create_masks, project_lengths_softmax
        for modname in ("mucon.masks", "mucon.models"):
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            for attr, obj in (("create_masks", create_masks), ("project_lengths_softmax", project_lengths_softmax)):
                if hasattr(mod, attr):
                    setattr(mod, attr, obj)
                    done.append((modname, attr))
    return done
