"""ctypes binding of oracle/oracle.c -- ORACLE / TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from . import poisson as _poisson

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_viterbi.restype = C.c_int
        _lib.orc_decode_video.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def block_scores(logp, fs):
    logp = np.ascontiguousarray(logp)
    T, Cc = logp.shape
    K = T // fs
    bs = np.empty((K, Cc), dtype=logp.dtype)
    fn = lib().orc_block_scores_f64 if logp.dtype == np.float64 else lib().orc_block_scores_f32
    fn(_p(logp), C.c_int64(T), C.c_int(Cc), C.c_int(fs), _p(bs))
    return bs


def poisson_rows(params_n3, fs, max_len):
    params = np.ascontiguousarray(params_n3, dtype=np.float64)
    N = params.shape[0]
    J = max_len // fs
    lf = _poisson.log_factorial_prefix(max_len - 1)
    rows = np.empty((N, J), dtype=np.float64)
    lib().orc_poisson_rows(_p(params), C.c_int(N), _p(lf), C.c_int(fs), C.c_int(max_len), _p(rows))
    return rows


def decode_video(logp, transcript, rows, fs, seg0_f32):
    """Whole video on the CPU in C.  Returns dict(score, labels, seg_blocks) or raises ValueError."""
    logp = np.ascontiguousarray(logp)
    T, Cc = logp.shape
    tr = np.ascontiguousarray(transcript, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    N, J = rows.shape
    score = C.c_double()
    seg = np.zeros(N, dtype=np.int32)
    labels = np.empty(T, dtype=np.int32)
    rc = lib().orc_decode_video(_p(logp), C.c_int(int(logp.dtype == np.float64)), C.c_int64(T), C.c_int(Cc),
                                _p(tr), C.c_int(N), _p(rows), C.c_int(J), C.c_int(fs), C.c_int(int(seg0_f32)),
                                C.byref(score), _p(seg), _p(labels))
    if rc != 0:
        raise ValueError("infeasible")
    return dict(score=score.value, labels=labels, seg_blocks=seg)


def viterbi(bs, transcript, rows, seg0_f32):
    bs = np.ascontiguousarray(bs)
    K, Cc = bs.shape
    tr = np.ascontiguousarray(transcript, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    N, J = rows.shape
    score = C.c_double()
    jf = C.c_int32()
    seg = np.zeros(N, dtype=np.int32)
    bp = np.zeros((K, N), dtype=np.uint16)
    rc = lib().orc_viterbi(_p(bs), C.c_int(int(bs.dtype == np.float64)), C.c_int64(K), C.c_int(Cc), _p(tr),
                           C.c_int(N), _p(rows), C.c_int(J), C.c_int(int(seg0_f32)),
                           C.byref(score), _p(seg), _p(bp), C.byref(jf))
    if rc != 0:
        raise ValueError("infeasible")
    return dict(score=score.value, seg_blocks=seg, bp=bp, jf=jf.value)


def masks(L, T, overlap, tmpl, align_corners):
    L = np.ascontiguousarray(L, dtype=np.float32)
    tmpl = np.ascontiguousarray(tmpl, dtype=np.float32)
    M = L.shape[0]
    out = np.empty((M, T), dtype=np.float32)
    L_out = np.empty(M, dtype=np.float32)
    lib().orc_masks(_p(L), C.c_int(M), C.c_int(T), C.c_float(overlap), _p(tmpl), C.c_int(tmpl.shape[0]),
                    C.c_int(int(align_corners)), _p(L_out), _p(out))
    return out, L_out
