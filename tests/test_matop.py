"""Matrix form of the interpolation (N4): PETScMatUtilities::constructPatchLevelSCInterpOp (PETScMatUtilities.cpp:783-1020).
CPU: the oracle's restatement applied to a field equals the oracle's (golden-pinned) interpolation funnel, rows sum to one.
GPU: libibk.so's rows equal the oracle's (same columns, values bit for bit) and J u equals interpolateVelocity."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit


def make_level(ndim, n, g, split=True):
    if split:
        boxes = [((0,) * ndim, (n // 2 - 1,) + (n - 1,) * (ndim - 1)), ((n // 2,) + (0,) * (ndim - 1), (n - 1,) * ndim)]
    else:
        boxes = [((0,) * ndim, (n - 1,) * ndim)]
    return orc.Level(ndim, (0,) * ndim, (n,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim, boxes, (g,) * ndim)


def global_dofs(level, n):
    """DOF number of side (i, axis) = axis * n^ndim + linear index of i mod n: one DOF per periodic class of sides"""
    ndim = level.ndim
    out = []
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        per_axis = []
        for a in range(ndim):
            gi = [np.mod(np.arange(pg.upper[d] - pg.lower[d] + 1 + (1 if d == a else 0) + 2 * pg.gcw[d]) + pg.lower[d] - pg.gcw[d], n)
                  for d in range(ndim)]
            mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
            lin = sum(mesh[d] * n ** d for d in range(ndim))
            per_axis.append((a * n ** ndim + lin).astype(np.int32))
        out.append(per_axis)
    return out


def markers(ndim, N):
    return np.stack([splitmix64_unit(40 + d, np.arange(N)) for d in range(ndim)], axis=1)


@pytest.mark.parametrize("ndim,fcn", [(2, "IB_4"), (3, "IB_4"), (2, "PIECEWISE_LINEAR"), (3, "PIECEWISE_LINEAR")])
def test_oracle_matrix_form_equals_the_funnel(ndim, fcn):
    n, N = 16, 500
    g = orc.min_ghost_width(fcn)
    level = make_level(ndim, n, g, split=False)
    X = markers(ndim, N)
    dof = global_dofs(level, n)
    cols, vals = orc.sc_interp_op(level, X, dof, fcn)
    assert cols.min() >= 0 and np.max(np.abs(vals.sum(axis=1) - 1.0)) < 1e-14
    u_vec = 2 * splitmix64_unit(9, np.arange(ndim * n ** ndim)) - 1
    U_mat = (vals * u_vec[cols]).sum(axis=1).reshape(N, ndim)
    pg = level.patch_geom(0)
    u = [np.ascontiguousarray(u_vec[dof[0][a]]) for a in range(ndim)]
    ref = orc.bin_level(level, X)
    lst = ref["patches"][0]
    ii = lst["all_idx"][lst["interior_mask"]]
    sh = lst["all_shift"].reshape(-1, ndim)[lst["interior_mask"]]
    U_ref = orc.side_interp(fcn, pg, u, X, ii, sh.reshape(-1))
    assert np.max(np.abs(U_mat - U_ref)) <= 1e-13 * np.max(np.abs(U_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("ndim,fcn", [(2, "IB_4"), (3, "IB_4"), (2, "PIECEWISE_LINEAR"), (3, "PIECEWISE_LINEAR")])
def test_device_rows_equal_the_oracle_and_apply_like_the_interpolation(ndim, fcn):
    from ibamr_b200 import api
    n, N = 32, 20000
    g = orc.min_ghost_width(fcn)
    level = make_level(ndim, n, g)
    X = markers(ndim, N)
    dof = global_dofs(level, n)
    ib = api.IBMethodB200(ndim, level.domain_lower, level.domain_upper(), level.x_lower, level.x_upper, level.periodic, level.boxes,
                          gcw=g, kernel_fcn=fcn, ctx=api.Context(0))
    try:
        ib.setPositions(X)
        ib.beginDataRedistribution()
        cols, vals = ib.constructInterpOp(dof, fcn)
        cols_ref, vals_ref = orc.sc_interp_op(level, X, dof, fcn)
        assert np.array_equal(cols, cols_ref)
        assert np.array_equal(vals, vals_ref)
        u_vec = 2 * splitmix64_unit(9, np.arange(ndim * n ** ndim)) - 1
        for p in range(len(level.boxes)):
            for a in range(ndim):
                ib.grid_upload("u", p, a, np.ascontiguousarray(u_vec[dof[p][a]]))
        ib.interpolateVelocity(fill_halo=False)
        U = ib.getLData("U")
        U_mat = (vals * u_vec[cols]).sum(axis=1).reshape(N, ndim)
        assert np.max(np.abs(U - U_mat)) <= 1e-12 * np.max(np.abs(U_mat))
    finally:
        ib.close()
