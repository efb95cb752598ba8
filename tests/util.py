import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("viterbi_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "viterbi_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"viterbi_{name}.npz"))
    g = {k: z[k] for k in z.files}
    off = g["tr_off"]
    g["transcripts"] = [g["tr"][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]
    g["fs"], g["max_len"] = int(g["fs"]), int(g["max_len"])
    g["segments"] = list(zip(g["seg_label"].tolist(), g["seg_length"].tolist()))
    # the float-promotion regime the fixture was minted under
    g["seg0_f32"] = bool(g["logp"].dtype == np.float32 and int(str(g["numpy_version"]).split(".")[0]) >= 2)
    return g


def same_score(a, b):
    a, b = np.float64(a), np.float64(b)
    return (a == b) or (np.isnan(a) and np.isnan(b))


def mask_atol(T, L):
    """Absolute tolerance for float32 masks: the template coordinate is evaluated in float32 from
    values as large as T/L, so its rounding error -- and a ramp value -- scales with T / min(L).  The
    floor is four float32 ulps of the coordinate's range (u up to 100, ulp 7.6e-6): the kernel's and
    torch's op orders differ by a few roundings of u on a ramp of slope 1."""
    return 3e-5 * max(1.0, float(T) / float(np.min(L)))
