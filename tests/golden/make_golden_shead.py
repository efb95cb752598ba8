"""Mints tests/golden/shead.npz from the UNMODIFIED reference MuCon.sequence_generation_forward
(/root/reference/src/mucon/models.py:585-745: BiLSTM encoder, attention decoder, transcript and length heads), on a
real MuCon model built with the reference's own default configuration (src/configs/mucon/default.py; yacs / fandak
stubbed as in make_golden_evaluator_flow.py, CfgNode = a plain attribute container).  Per video: a random encoded
sequence z [1, Tz, 128], the teacher-forced run (transcript log-probabilities and length logits of every step) and
the greedy run (eval mode, teacher forcing off: stops at EOS).  The state_dict of the s-head parameters is stored so
the test does not depend on RNG reproducibility."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_evaluator_flow as flow  # noqa: E402

flow.install_stubs()


class CN(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)


sys.modules["yacs.config"].CfgNode = CN
core_config = types.ModuleType("core.config")
core_config.dataset_cfg, core_config.system_cfg = CN(), CN()
sys.modules["core.config"] = core_config
sys.modules["fandak"].Model = type("Model", (nn.Module,), {"__init__": lambda self, cfg=None: (nn.Module.__init__(self), setattr(self, "cfg", cfg))[0]})

from configs.mucon.default import get_cfg_defaults  # noqa: E402
from mucon.models import MuCon  # noqa: E402

C = 48


def main():
    torch.manual_seed(0)
    cfg = get_cfg_defaults()
    model = MuCon(cfg, input_feature_size=2048, num_classes=C, max_decoding_steps=31).eval()
    out = {"torch_version": torch.__version__}
    for k, v in model.state_dict().items():
        if k.startswith("fs_"):
            out["w." + k] = v.numpy()
    rng = np.random.default_rng(4)
    cases = [(37, 5), (125, 8), (240, 3), (9, 2)]
    with torch.no_grad():
        for i, (Tz, N) in enumerate(cases):
            z = torch.from_numpy(np.maximum(rng.standard_normal((1, Tz, 128)), 0).astype(np.float32))   # post-ReLU like
            tr = rng.integers(0, C, N)
            tf_in = torch.from_numpy(np.concatenate([[C + 1], tr])).long()      # SOS + transcript (general_dataset.py)
            tf_tgt = torch.from_numpy(np.concatenate([tr, [C]])).long()         # transcript + EOS
            model.set_teacher_forcing(True)
            pt, pl = model.sequence_generation_forward(z, N + 1, tf_in, tf_tgt)
            out[f"c{i}_z"], out[f"c{i}_tf_in"] = z[0].numpy(), tf_in.numpy()
            out[f"c{i}_tf_logp"] = torch.cat(pt, 0).numpy()
            out[f"c{i}_tf_len"] = torch.stack(pl).numpy()
            model.set_teacher_forcing(False)
            pt, pl = model.sequence_generation_forward(z, N + 1, tf_in, tf_tgt)
            out[f"c{i}_greedy_logp"] = torch.cat(pt, 0).numpy()
            out[f"c{i}_greedy_len"] = torch.stack(pl).numpy()
            print(i, Tz, N, "greedy steps", len(pt), [int(p.argmax()) for p in pt][:8])
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "shead.npz"), **out)
    print("wrote shead.npz")


if __name__ == "__main__":
    main()
