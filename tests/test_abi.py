"""CPU: the C-ABI library loads and exports every symbol include/mucon_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from mucon_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "mucon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mucon_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mucon_b200.h but not exported"


def test_loader_symbol_list_matches_header(built_lib):
    from mucon_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared()
    assert _lib.lib().mucon_abi_version() == 2


def test_struct_layout_matches_header():
    from mucon_b200 import _lib
    # 12 int32 + 18 pointers + peer_delta[8] (int64)
    assert ctypes.sizeof(_lib.ViterbiBatch) == 12 * 4 + 18 * 8 + 8 * 8


def test_argument_validation_without_gpu(built_lib):
    from mucon_b200 import _lib
    l = _lib.lib()
    assert l.mucon_viterbi_decode(None, None) == -1
    assert l.mucon_viterbi_blockscores(None, 0, None, None, None, 1, 48, 30, None, None) == -1
    assert l.mucon_strerror(-2).decode().startswith("shape")


def test_host_helpers(built_lib):
    import numpy as np
    from mucon_b200 import _lib
    from mucon_b200.length_model import log_factorial_prefix, poisson_params
    l = _lib.lib()
    means = np.array([1.0, 333.25, 57.5, 0.7, 4500.5])
    out = np.zeros((5, 3))
    assert l.mucon_poisson_params_h(means.ctypes.data_as(ctypes.c_void_p), 5, out.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = poisson_params(means)
    assert np.allclose(out, ref, rtol=1e-15, atol=0)  # libm vs numpy ln(): <= 1 ulp
    lf = np.zeros(67)
    assert l.mucon_logfact_h(30, 2000, lf.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.allclose(lf, log_factorial_prefix(1999)[np.arange(67) * 30], rtol=1e-15, atol=0)


def _pack_h(lib, N, order, max_N, lanes=0, fs=30, max_len=2000):
    N = np.ascontiguousarray(N, dtype=np.int32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    wu = np.full(max(len(N), 1) * 16, -1, dtype=np.int32)
    n_cta, wpc, lanes_out = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
    rc = lib.mucon_viterbi_pack_h(N.ctypes.data_as(ctypes.c_void_p), order.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int(len(N)), ctypes.c_int(max_N), ctypes.c_int(fs), ctypes.c_int(max_len),
                                  ctypes.c_int(lanes), wu.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_cta),
                                  ctypes.byref(wpc), ctypes.byref(lanes_out))
    return rc, wu[:n_cta.value * max(wpc.value, 1)].reshape(n_cta.value, max(wpc.value, 1)), wpc.value, lanes_out.value


def test_pack_h_assigns_every_unit_a_contiguous_run_of_warps(built_lib):
    """Host packing of units into DP bins (no GPU needed): every unit appears exactly once, on
    ceil((N-1)/segments_per_warp) consecutive warps of one bin; bins hold 4, 8 or 16 warps."""
    lib = ctypes.CDLL(built_lib)
    rng = np.random.default_rng(0)
    for trial in range(20):
        U = int(rng.integers(1, 200))
        max_N = int(rng.integers(1, 60))
        N = rng.integers(1, max_N + 1, U)
        order = np.argsort(-N, kind="stable")
        rc, wu, wpc, lanes = _pack_h(lib, N, order, int(N.max()))
        assert rc == 0 and wpc in (4, 8, 16) and lanes == 8  # J = 66 -> 8 lanes per segment
        spw = 32 // lanes
        seen = {}
        for b in range(wu.shape[0]):
            row = wu[b]
            for u in set(row[row >= 0].tolist()):
                pos = np.nonzero(row == u)[0]
                assert u not in seen and np.array_equal(pos, np.arange(pos[0], pos[0] + len(pos)))
                assert len(pos) == max(1, -(-(int(N[u]) - 1) // spw))
                seen[u] = b
        assert sorted(seen) == list(range(U))
    # shapes the register kernels do not cover are refused, not mis-packed
    assert _pack_h(lib, [3], [0], 3, max_len=6000)[0] == -2      # J = 200 > 128
    assert _pack_h(lib, [70], [0], 70)[0] == -2                  # N > 65


def test_pack_lanes_h_packs_units_side_by_side(built_lib):
    """Lane-per-segment packing: a unit takes max(1, N-1) consecutive lanes of one warp."""
    lib = ctypes.CDLL(built_lib)
    rng = np.random.default_rng(1)
    for trial in range(20):
        U = int(rng.integers(1, 300))
        N = rng.integers(1, 34, U).astype(np.int32)
        order = np.argsort(-N, kind="stable").astype(np.int32)
        lu = np.full(U * 32, -1, dtype=np.int32)
        nw = ctypes.c_int32(0)
        assert lib.mucon_viterbi_pack_lanes_h(N.ctypes.data_as(ctypes.c_void_p), order.ctypes.data_as(ctypes.c_void_p),
                                              ctypes.c_int(U), lu.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nw)) == 0
        lanes = lu[:nw.value * 32].reshape(nw.value, 32)
        seen = set()
        for w in range(nw.value):
            row = lanes[w]
            for u in set(row[row >= 0].tolist()):
                pos = np.nonzero(row == u)[0]
                assert u not in seen and np.array_equal(pos, np.arange(pos[0], pos[0] + len(pos)))
                assert len(pos) == max(1, int(N[u]) - 1)
                seen.add(u)
        assert seen == set(range(U))
        assert nw.value <= U and (lanes >= 0).sum() == np.maximum(N - 1, 1).sum()
    big = np.array([40], dtype=np.int32)
    assert lib.mucon_viterbi_pack_lanes_h(big.ctypes.data_as(ctypes.c_void_p), None, ctypes.c_int(1),
                                          lu.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nw)) == -2
