"""CPU: the C-ABI library loads and exports every symbol include/mucon_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

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
    assert _lib.lib().mucon_abi_version() == 1


def test_struct_layout_matches_header():
    from mucon_b200 import _lib
    # 12 int32 + 18 pointers
    assert ctypes.sizeof(_lib.ViterbiBatch) == 12 * 4 + 18 * 8


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
