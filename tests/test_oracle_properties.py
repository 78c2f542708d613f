"""Properties that tie the oracle's SPREAD to its INTERPOLATION (CPU only).

The reference's fixtures pin interpolation values for every kernel (tests/test_oracle_golden.py) but spread values only for
PIECEWISE_LINEAR (ghost_accumulation_01).  SURVEY.md 8(c): the other kernels' spreading is pinned transitively, through the
discrete adjointness  <S F, u> h^d = <F, J u>  of the two operators (same stencils, same weights), the zeroth moment
(sum of the spread field times h^d = sum of the forces) and the reproduction of constants / linear fields by J.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit

KERNELS = sorted(orc.KERNELS)
# kernels whose interpolation reproduces linear fields exactly (first moment condition); PIECEWISE_CONSTANT and
# DISCONTINUOUS_LINEAR (constant across the component's axis) do not
LINEAR = [k for k in KERNELS if k not in ("PIECEWISE_CONSTANT", "DISCONTINUOUS_LINEAR")]


def _geom(ndim, kernel, n=12):
    g = orc.min_ghost_width(kernel)
    return orc.PatchGeom((0,) * ndim, (n - 1,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1.0 / n,) * ndim, (g,) * ndim), n


def _markers(ndim, N, seed):
    # anywhere in the patch: with the minimum ghost width no stencil is clipped
    return np.stack([0.002 + 0.996 * splitmix64_unit(seed + d, np.arange(N)) for d in range(ndim)], axis=1)


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", KERNELS)
def test_side_spread_is_the_adjoint_of_side_interpolation(kernel, ndim):
    pg, n = _geom(ndim, kernel)
    N = 200
    X = _markers(ndim, N, 100)
    F = np.stack([2 * splitmix64_unit(200 + d, np.arange(N)) - 1 for d in range(ndim)], axis=1)
    u = [2 * splitmix64_unit(300 + a, np.arange(int(np.prod(pg.side_shape(a))))).reshape(pg.side_shape(a)) - 1 for a in range(ndim)]
    U = orc.side_interp_positions(kernel, pg, u, X)
    f = [np.zeros(pg.side_shape(a)) for a in range(ndim)]
    orc.side_spread_positions(kernel, pg, f, X, F)
    h = (1.0 / n) ** ndim
    lhs = sum(float(np.sum(f[a] * u[a])) for a in range(ndim)) * h
    rhs = float(np.sum(F * U))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(rhs), 1.0)
    # zeroth moment per component: everything that was spread is on the grid (ghosts included)
    for a in range(ndim):
        assert abs(float(np.sum(f[a])) * h - float(np.sum(F[:, a]))) <= 1e-12 * N


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", KERNELS)
def test_cell_spread_is_the_adjoint_of_cell_interpolation(kernel, ndim):
    pg, n = _geom(ndim, kernel)
    N, depth = 150, 2
    X = _markers(ndim, N, 400)
    Q = np.stack([2 * splitmix64_unit(500 + d, np.arange(N)) - 1 for d in range(depth)], axis=1)
    shape = (depth,) + tuple(pg.cell_shape())
    u = (2 * splitmix64_unit(600, np.arange(int(np.prod(shape)))) - 1).reshape(shape)
    V = orc.cell_interp_positions(kernel, pg, u, depth, X)
    q = np.zeros(shape)
    orc.cell_spread_positions(kernel, pg, q, depth, X, Q)
    h = (1.0 / n) ** ndim
    assert abs(float(np.sum(q * u)) * h - float(np.sum(Q * V))) <= 1e-12 * max(abs(float(np.sum(Q * V))), 1.0)


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", KERNELS)
def test_interpolation_reproduces_constants_and_linear_fields(kernel, ndim):
    pg, n = _geom(ndim, kernel)
    X = _markers(ndim, 120, 700)
    coef = [0.7, -1.3, 2.1][:ndim]
    const = [np.full(pg.side_shape(a), 3.25) for a in range(ndim)]
    U = orc.side_interp_positions(kernel, pg, const, X)
    # (BSPLINE_6's degree-5 polynomial in r = |x| + 3 cancels terms of 1e4: its weights sum to 1 within a few 1e-13 only)
    assert np.max(np.abs(U - 3.25)) <= (1e-11 if "6" in kernel and "BSPLINE" in kernel else 1e-12)
    if kernel in LINEAR:
        lin = []
        for a in range(ndim):
            c = pg.side_coords(a)
            lin.append(np.ascontiguousarray(1.5 + sum(coef[d] * c[d] for d in range(ndim))))
        U = orc.side_interp_positions(kernel, pg, lin, X)
        exact = 1.5 + X @ np.asarray(coef)
        assert np.max(np.abs(U - exact[:, None])) <= 1e-11


def test_wall_fold_back_is_the_adjoint_of_the_robin_extrapolation():
    """oracle.fold_walls (the adjoint_op = 1 branch of fortran/cartphysbdryop3d.f.m4:78-168, 787-905) against the forward
    branch of the same routines restated here: <v, E u> over the ghosted array = <E^T v, u> over the interior."""
    n, g = 8, 3
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 0), [((0,) * 3, (n - 1,) * 3)], (g,) * 3)
    pg = level.patch_geom(0)
    rng = np.random.default_rng(0)
    for bval in (0.0, 0.3):
        a, b = np.ones((3, 2, 3)), np.full((3, 2, 3), bval)
        h = level.dx[2]

        def extrap(u):  # homogeneous ghost-cell extrapolation, adjoint_op = 0
            u = [x.copy() for x in u]
            for side in (0, 1):
                sgn = -1 if side == 0 else 1
                for comp in range(3):
                    A = u[comp]
                    if comp == 2:
                        ib = g if side == 0 else g + n
                        if bval == 0.0:
                            A[ib] = 0.0
                        for i in range(1, g + 1):
                            if bval == 0.0:
                                A[ib + sgn * i] = -A[ib - sgn * i] + 2.0 * A[ib]
                            else:
                                A[ib + sgn * i] = A[ib - sgn * i] + (-a[2, side, comp] * (2.0 * i) * h / bval) * A[ib]
                    else:
                        ii = g if side == 0 else g + n - 1
                        for i in range(g):
                            nn = 1 + 2 * i
                            fi = -(a[2, side, comp] * nn * h - 2 * bval) / (a[2, side, comp] * nn * h + 2 * bval)
                            A[ii + sgn * (1 + i)] = fi * A[ii - sgn * i]
            return u

        def transverse_interior(c):
            m = np.zeros(pg.side_shape(c), bool)
            sl = [slice(None)] * 3
            for e in (0, 1):
                sl[2 - e] = slice(g, g + n + (1 if e == c else 0))
            m[tuple(sl)] = True
            return m

        u = [rng.standard_normal(pg.side_shape(c)) * transverse_interior(c) for c in range(3)]
        v = [rng.standard_normal(pg.side_shape(c)) * transverse_interior(c) for c in range(3)]
        if bval == 0.0:  # Dirichlet pins the boundary face: it is not a degree of freedom of u
            for side_idx in (g, g + n):
                u[2][side_idx] = 0.0
        Eu = extrap(u)
        Fv = orc.fold_walls(level, 0, [x.copy() for x in v], a, b)

        def zin(c, arr):
            return arr[g:g + n + (1 if c == 2 else 0)]
        lhs = sum(np.vdot(v[c], Eu[c]) for c in range(3))
        rhs = sum(np.vdot(zin(c, Fv[c]), zin(c, u[c])) for c in range(3))
        assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)
