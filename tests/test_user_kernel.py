"""The USER_DEFINED kernel (N4): LEInteractor::s_kernel_fcn + userDefinedInterpolate / userDefinedSpread
(LEInteractor.cpp:2019-2020, 6128-6382).  CPU: the oracle's restatement with the reference's default function (the 4-point
function) agrees with the golden-pinned IB_4 funnel.  GPU: libibk.so with a callback registered through ibk_set_user_kernel
against the oracle -- interpolation and spread bit for bit (the spread adds in the reference's serial order) -- at the
funnel seam, and against the built-in IB_4 at the patch seam."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit


def cosine_kernel(r):  # a smooth 4-point function that is not one of the built-in ones
    r = abs(r)
    return 0.25 * (1.0 + math.cos(0.5 * math.pi * r)) if r < 2.0 else 0.0


def hat3(r):  # an odd stencil: the quadratic B-spline
    r = abs(r)
    if r < 0.5:
        return 0.75 - r * r
    if r < 1.5:
        return 0.5 * (1.5 - r) * (1.5 - r)
    return 0.0


def case(ndim, n, g, N, depth, seed=0):
    dx = (1.0 / n,) * ndim
    ilower, iupper = (0,) * ndim, (n - 1,) * ndim
    shape = (depth,) + tuple(n + 2 * g for _ in range(ndim))
    cnt = int(np.prod(shape))
    u = (2 * splitmix64_unit(50 + seed, np.arange(cnt)) - 1).reshape(shape)
    X = np.stack([-0.03 + 1.06 * splitmix64_unit(60 + d + seed, np.arange(N)) for d in range(ndim)], axis=1)  # some beyond the patch
    V = np.stack([2 * splitmix64_unit(70 + d + seed, np.arange(N)) - 1 for d in range(depth)], axis=1)
    idx = np.arange(N - 1, -1, -1, dtype=np.int32)[: N - 5]  # a list that is not the identity
    sh = 1e-3 * (2 * splitmix64_unit(80 + seed, np.arange(idx.size * ndim)) - 1)
    return dx, ilower, iupper, (g,) * ndim, u, X, V, idx, sh


@pytest.mark.parametrize("ndim", [2, 3])
def test_oracle_user_defined_with_the_default_function_is_ib_4(ndim):
    n, g, N, depth = 12, 3, 150, 2
    dx, ilo, iup, nug, u, X, V, idx, _ = case(ndim, n, g, N, depth)
    X = 0.05 + 0.9 * (X + 0.03) / 1.06  # interior stencils only: the clipping rules of the two routines differ
    xl = (0.0,) * ndim
    U1 = orc.user_interp_raw(orc.ib4_kernel_fcn, 4, ndim, dx, xl, depth, ilo, iup, nug, u, idx, None, X, np.zeros((N, depth)))
    U2 = orc.interp_raw("IB_4", ndim, dx, xl, depth, ilo, iup, nug, u, idx, np.zeros(idx.size * ndim), X, np.zeros((N, depth)))
    assert np.max(np.abs(U1 - U2.reshape(N, depth))) < 1e-13
    f1 = orc.user_spread_raw(orc.ib4_kernel_fcn, 4, ndim, dx, xl, depth, idx, None, X, V, ilo, iup, nug, np.zeros_like(u))
    f2 = orc.spread_raw("IB_4", ndim, dx, xl, depth, idx, np.zeros(idx.size * ndim), X, V, ilo, iup, nug, np.zeros_like(u))
    assert np.max(np.abs(f1 - f2.reshape(f1.shape))) < 1e-13 * np.max(np.abs(f1))


@pytest.mark.gpu
@pytest.mark.parametrize("ndim,fn,stencil", [(2, cosine_kernel, 4), (3, cosine_kernel, 4), (3, hat3, 3), (2, orc.ib4_kernel_fcn, 4)])
def test_device_user_defined_at_the_funnel_seam_bit_exact(ndim, fn, stencil):
    from ibamr_b200 import api
    ctx = api.default_context()
    n, g, N, depth = 16, 3, 400 if ndim == 3 else 1500, 2
    dx, ilo, iup, nug, u, X, V, idx, sh = case(ndim, n, g, N, depth)
    xl, xu = (0.0,) * ndim, (1.0,) * ndim
    cb = C.CFUNCTYPE(C.c_double, C.c_double)(lambda r: float(fn(r)))
    assert ctx.lib.ibk_set_user_kernel(cb, stencil) == 0
    try:
        assert ctx.lib.ibk_get_stencil_size(b"USER_DEFINED") == stencil
        U0 = 7.0 * np.ones((N, depth))
        U = api.raw_interp_host("USER_DEFINED", ndim, dx, xl, xu, depth, ilo, iup, nug, u, idx, sh, X, U0.copy())
        U_ref = orc.user_interp_raw(fn, stencil, ndim, dx, xl, depth, ilo, iup, nug, u, idx, sh, X, U0.copy())
        assert np.array_equal(U.reshape(N, depth), U_ref)  # unlisted markers keep their values, listed ones agree bit for bit
        f0 = 0.5 * np.ones_like(u)
        f = api.raw_spread_host("USER_DEFINED", ndim, dx, xl, xu, depth, idx, sh, X, V, ilo, iup, nug, f0.copy())
        f_ref = orc.user_spread_raw(fn, stencil, ndim, dx, xl, depth, idx, sh, X, V, ilo, iup, nug, f0.copy())
        assert np.array_equal(f.reshape(f_ref.shape), f_ref)
    finally:
        ctx.lib.ibk_set_user_kernel(C.cast(None, C.CFUNCTYPE(C.c_double, C.c_double)), 0)


@pytest.mark.gpu
def test_device_user_defined_at_the_patch_seam_equals_builtin_ib_4():
    """Side data, position-only overloads: with the reference's default function registered, "USER_DEFINED" and "IB_4" agree."""
    from ibamr_b200 import api
    n, N = 24, 5000
    box = api.Box((0, 0, 0), (n - 1,) * 3)
    pd = api.Patch(box, (0.0,) * 3, (1.0,) * 3, (1.0 / n,) * 3)
    ctx = api.default_context()
    ctx.lib.ibk_set_user_kernel(C.cast(None, C.CFUNCTYPE(C.c_double, C.c_double)), 0)  # the default: the 4-point function
    X = np.stack([0.1 + 0.8 * splitmix64_unit(30 + d, np.arange(N)) for d in range(3)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    res = {}
    for k in ("IB_4", "USER_DEFINED"):
        u = api.SideData(box, 1, 3)
        rng = np.random.default_rng(3)
        for a in range(3):
            u.arrays[a][...] = rng.standard_normal(u.arrays[a].shape)
        U = np.zeros((N, 3))
        api.LEInteractor.interpolate(U, 3, X, 3, u, pd, box, k)
        f = api.SideData(box, 1, 3)
        api.LEInteractor.spread(f, F, 3, X, 3, pd, box, k)
        res[k] = (U, [a.copy() for a in f.arrays])
    assert np.max(np.abs(res["IB_4"][0] - res["USER_DEFINED"][0])) < 1e-13 * np.max(np.abs(res["IB_4"][0]))
    for a in range(3):
        assert np.max(np.abs(res["IB_4"][1][a] - res["USER_DEFINED"][1][a])) < 1e-13 * np.max(np.abs(res["IB_4"][1][a]))
