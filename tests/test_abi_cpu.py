"""CPU-side checks of the drop-in boundary: libibk.so loads, exports every symbol include/ibk.h
declares, answers the LEInteractor static queries, and refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ibamr_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ibk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ibk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from ibamr_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (ibk_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/ibk.h but not exported by libibk.so: {missing}"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SYMBOLS) == declared


def test_kernel_queries_match_leinteractor(lib):
    # LEInteractor::getStencilSize / getMinimumGhostWidth (LEInteractor.cpp:2052-2114)
    expect = {"IB_4": (4, 3), "IB_6": (6, 4), "BSPLINE_3": (4, 3), "BSPLINE_4": (4, 3), "PIECEWISE_LINEAR": (2, 2), "IB_3": (4, 3),
              "BSPLINE_5": (6, 4), "BSPLINE_6": (6, 4), "PIECEWISE_CUBIC": (4, 3), "IB_5": (6, 4), "PIECEWISE_CONSTANT": (1, 1),
              "COMPOSITE_BSPLINE_32": (4, 3), "COMPOSITE_BSPLINE_23": (4, 3), "COMPOSITE_BSPLINE_43": (4, 3), "COMPOSITE_BSPLINE_34": (4, 3),
              "COMPOSITE_BSPLINE_54": (5, 3), "COMPOSITE_BSPLINE_45": (5, 3), "COMPOSITE_BSPLINE_65": (6, 4), "COMPOSITE_BSPLINE_56": (6, 4),
              "DISCONTINUOUS_LINEAR": (2, 2), "IB_4_W8": (8, 5)}
    for name, (sz, g) in expect.items():
        assert lib.ibk_is_known_kernel(name.encode()) == 1
        assert lib.ibk_get_stencil_size(name.encode()) == sz
        assert lib.ibk_get_minimum_ghost_width(name.encode()) == g
    # "USER_DEFINED": the statics of LEInteractor (defaults LEInteractor.cpp:2019-2020: the 4-point function, stencil 4)
    assert lib.ibk_is_known_kernel(b"USER_DEFINED") == 1 and lib.ibk_get_stencil_size(b"USER_DEFINED") == 4
    cb = C.CFUNCTYPE(C.c_double, C.c_double)(lambda r: max(0.0, 1.0 - abs(r)))
    assert lib.ibk_set_user_kernel(cb, 2) == 0 and lib.ibk_get_stencil_size(b"USER_DEFINED") == 2
    assert lib.ibk_get_minimum_ghost_width(b"USER_DEFINED") == 2
    assert lib.ibk_set_user_kernel(cb, 0) == -1  # IBK_ERR_INVALID
    assert lib.ibk_set_user_kernel(C.cast(None, C.CFUNCTYPE(C.c_double, C.c_double)), 0) == 0  # back to the default
    assert lib.ibk_get_stencil_size(b"USER_DEFINED") == 4
    assert lib.ibk_is_known_kernel(b"IB_7") == 0
    assert lib.ibk_get_stencil_size(b"NOPE") == -3  # IBK_ERR_UNKNOWN_KERNEL
    assert lib.ibk_kernel_from_string(b"IB_4") == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is only observable on a CPU-only box")
    h = C.c_void_p()
    assert lib.ibk_ctx_create(0, C.byref(h)) == -2  # IBK_ERR_CUDA
    assert not h.value
    from ibamr_b200 import api
    with pytest.raises(api.IBKError):
        api.Context(0)
