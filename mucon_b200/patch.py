"""patch_reference(): installs the GPU implementations into an importable reference tree
(`core.viterbi.viterbi.Viterbi`, `mucon.masks.create_masks`, `core.modules.temporal.WaveNetBlock`, ...) without
editing its sources; unpatch_reference() puts the originals back.  patch_model() converts an already constructed
reference model in place (its `ft` block becomes this package's WaveNetBlock with the same weights)."""
import importlib

_ORIGINALS = []


def _swap(modname, attr, obj, done):
    try:
        mod = importlib.import_module(modname)
    except Exception:
        return
    if hasattr(mod, attr):
        _ORIGINALS.append((mod, attr, getattr(mod, attr)))
        setattr(mod, attr, obj)
        done.append((modname, attr))


def patch_reference(viterbi=True, masks=True, backbone=True):
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
            _swap(modname, attr, obj, done)
    if masks:
        from .masks import create_masks, project_lengths_softmax
        for modname in ("mucon.masks", "mucon.models"):
            for attr, obj in (("create_masks", create_masks), ("project_lengths_softmax", project_lengths_softmax)):
                _swap(modname, attr, obj, done)
    if backbone:
        # models created after this call build the GPU blocks (same constructor signatures and parameter names:
        # core/modules/temporal.py:77-126,150-186, selected at mucon/models.py:160-186); inference (eval) forward
        from . import temporal
        for modname in ("core.modules.temporal", "mucon.models"):
            for attr in ("WaveNetBlock", "MSTCNPPFirstStage", "NoFt"):
                _swap(modname, attr, getattr(temporal, attr), done)
    return done


def unpatch_reference():
    """Restores everything patch_reference() replaced (in reverse order)."""
    while _ORIGINALS:
        mod, attr, obj = _ORIGINALS.pop()
        setattr(mod, attr, obj)


def patch_model(model):
    """An existing reference MuCon model: replace its temporal block `ft` (a reference WaveNetBlock /
    MSTCNPPFirstStage / NoFt) by this package's block of the same configuration and weights.  Returns the model."""
    from . import temporal
    ft = model.ft
    name = type(ft).__name__
    if name == "WaveNetBlock":
        new = temporal.WaveNetBlock(ft.in_channels, stages=list(ft.stages), out_dims=ft.out_dims,
                                    kernel_size=ft.kernel_size, pooling=ft.pooling, pooling_layers=list(ft.pooling_layers),
                                    pooling_type=ft.pooling_type, dropout_rate=ft.dropout_rate, leaky=ft.leaky)
    elif name == "MSTCNPPFirstStage":
        new = temporal.MSTCNPPFirstStage(ft.num_layers, ft.num_f_maps, ft.input_dim, ft.output_dim,
                                         pooling_layers=list(ft.pooling_layers))
    elif name == "NoFt":
        new = temporal.NoFt(ft.in_chnnels, ft.out_dims, ft.kernel_size)
    else:
        raise TypeError(f"unknown temporal block {name}")
    new.load_state_dict(ft.state_dict())
    new.to(next(ft.parameters()).device)
    new.train(ft.training)
    model.ft = new
    return model
