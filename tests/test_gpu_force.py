"""N1 on the GPU: Lagrangian forces, column algebra and one whole explicit IB step with X, U, F
resident on the device, against the oracle's restatement of IBStandardForceGen / IBMethod
(oracle/le_force.c, src/IB/IBStandardForceGen.cpp:813-1299, src/IB/IBMethod.cpp:714-858)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def api():
    from ibamr_b200 import api as _api
    _api.default_context()
    return _api


def _u(seed, n, lo=0.0, hi=1.0):
    return lo + (hi - lo) * splitmix64_unit(seed, np.arange(n))


def _random_elements(ndim, N, seed):
    ns, nb, nt = 3 * N, 2 * N, N // 3
    m = (_u(seed, ns) * N).astype(np.int32)
    s = (m + 1 + (_u(seed + 1, ns) * (N - 1)).astype(np.int32)) % N  # never equal to the master
    springs = (m, s.astype(np.int32), _u(seed + 2, ns, 0.5, 3.0), _u(seed + 3, ns, 0.0, 0.2))
    c = (_u(seed + 4, nb) * N).astype(np.int32)
    nx = (c + 1 + (_u(seed + 5, nb) * (N - 1)).astype(np.int32)) % N
    pv = (c + 1 + (_u(seed + 6, nb) * (N - 1)).astype(np.int32)) % N
    beams = (c, nx.astype(np.int32), pv.astype(np.int32), _u(seed + 7, nb, 0.1, 2.0),
             np.stack([_u(seed + 8 + d, nb, -0.1, 0.1) for d in range(ndim)], axis=1))
    ti = np.unique((_u(seed + 12, nt) * N).astype(np.int32))
    targets = (ti, _u(seed + 13, len(ti), 1.0, 5.0), _u(seed + 14, len(ti), 0.0, 1.0),
               np.stack([_u(seed + 15 + d, len(ti)) for d in range(ndim)], axis=1))
    return springs, beams, targets


@pytest.mark.parametrize("ndim", [2, 3])
def test_lagrangian_force_vs_oracle(api, ndim):
    """Random springs / beams / target points (several elements per node, nodes without any), evaluated in the
    upload order and again after a re-bin permuted the storage order: same bits both times, and equal to the
    serial oracle to rounding (the per-node summation order is the reference's)."""
    n, N = 16, 3000
    boxes = [((0,) * ndim, (n - 1,) * ndim)]
    ib = api.IBMethodB200(ndim, (0,) * ndim, (n - 1,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim, boxes, kernel_fcn="IB_4")
    X = np.stack([_u(200 + d, N) for d in range(ndim)], axis=1)
    U = np.stack([_u(210 + d, N, -1.0, 1.0) for d in range(ndim)], axis=1)
    springs, beams, targets = _random_elements(ndim, N, 300)
    ib.setPositions(X)
    ib.setLData("U", U)
    ib.setSprings(*springs)
    ib.setBeams(beams[0], beams[1], beams[2], beams[3], beams[4])
    ib.setTargetPoints(*targets)
    F_ref = orc.lagrangian_force(ndim, X, U, springs=springs, beams=beams, targets=targets)
    ib.computeLagrangianForce()
    F0 = ib.getLData("F")
    ib.beginDataRedistribution()
    ib.computeLagrangianForce()
    F1 = ib.getLData("F")
    assert np.array_equal(F0, F1)
    scale = np.max(np.abs(F_ref))
    assert np.max(np.abs(F0 - F_ref)) <= 1e-14 * scale
    # each group alone (the others cleared)
    for kw in ({"springs": springs}, {"beams": beams}, {"targets": targets}):
        ib.clearForces()
        if "springs" in kw:
            ib.setSprings(*springs)
        if "beams" in kw:
            ib.setBeams(beams[0], beams[1], beams[2], beams[3], beams[4])
        if "targets" in kw:
            ib.setTargetPoints(*targets)
        ib.computeLagrangianForce()
        Fr = orc.lagrangian_force(ndim, X, U, **kw)
        assert np.max(np.abs(ib.getLData("F") - Fr)) <= 1e-14 * np.max(np.abs(Fr))
    ib.close()


def test_column_algebra_and_anchor_rows(api):
    ndim, n, N = 3, 16, 5000
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1,) * 3, [((0,) * 3, (n - 1,) * 3)], kernel_fcn="IB_4")
    X = np.stack([_u(400 + d, N) for d in range(3)], axis=1)
    U = np.stack([_u(410 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ib.setPositions(X)
    ib.setLData("U", U)
    ib.beginDataRedistribution()  # storage order != Lagrangian order from here on
    ib.preprocessIntegrateData()
    assert np.array_equal(ib.getLData("X_current"), X)
    dt = 0.0371
    ib.forwardEulerStep(dt)
    Xn = X + dt * U  # VecWAXPY: one multiply, one add per entry
    assert np.array_equal(ib.getLData("X_new"), Xn)
    assert np.array_equal(ib.getLData("X"), 0.5 * X + 0.5 * Xn)
    anchors = np.array([0, 17, N - 1, 17], dtype=np.int32)
    ib.resetAnchorPointValues("U", anchors)
    U2 = U.copy()
    U2[anchors] = 0.0
    assert np.array_equal(ib.getLData("U"), U2)
    ds = _u(420, N, 0.5, 2.0)
    ib.scaleRows("aux", "U", ds)  # F * ds of LDataManager::spread, here on U
    assert np.array_equal(ib.getLData("aux"), U2 * ds[:, None])
    ib.postprocessIntegrateData()
    assert np.array_equal(ib.getLData("X"), Xn)
    with pytest.raises(api.IBKError):
        ib.spreadForce()  # positions moved: a re-bin is required first
    ib.close()


def test_one_explicit_ib_step_resident_on_the_device(api):
    """examples/IB/explicit/ex1 (curve2d_64.vertex/.spring, IB_4, 64^2 periodic): one midpoint-rule step of
    IBExplicitHierarchyIntegrator with the marker data never leaving the device:
        U = J u^n at X^n; X^{n+1,*} = X^n + dt U; X^{n+1/2} = (X^n + X^{n+1,*})/2      (forwardEulerStep)
        F = F(X^{n+1/2}); f = S F                                                        (computeLagrangianForce, spreadForce)
        u^{n+1/2} = u^n + c f  (stand-in for the fluid solve, on the host for both sides)
        U = J u^{n+1/2} at X^{n+1/2}; X^{n+1} = X^n + dt U                               (midpointStep)
    against the oracle's CPU restatement of every stage."""
    ndim, n, kernel, dt = 2, 64, "IB_4", 2.5e-3
    init = api.IBStandardInitializer(2, [os.path.join(GOLD, "curve2d_64")])
    g = orc.min_ghost_width(kernel)
    level = orc.Level(2, (0, 0), (n, n), (0.0, 0.0), (1.0, 1.0), (1, 1), [((0, 0), (n - 1, n - 1))], (g, g))
    pg = level.patch_geom(0)
    ib = api.IBMethodB200(2, (0, 0), (n - 1, n - 1), (0.0, 0.0), (1.0, 1.0), (1, 1), level.boxes, gcw=g, kernel_fcn=kernel)
    X0 = init.register(ib)
    m, s, k, r, _ = init.springs[0]

    def u_field(scale):
        out = []
        for a in range(2):
            c = pg.side_coords(a)
            out.append(np.ascontiguousarray(scale * (np.sin(2 * np.pi * c[0]) * np.cos(2 * np.pi * c[1]) if a == 0 else
                                                     -np.cos(2 * np.pi * c[0]) * np.sin(2 * np.pi * c[1]))))
        return out

    def oracle_interp(u, X):
        ref = orc.bin_level(level, X)
        lst = ref["patches"][0]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, 2)[lst["interior_mask"]]
        U = np.zeros_like(X)
        orc.side_interp(kernel, pg, u, X, ii, sh.reshape(-1), U)
        return U

    def oracle_spread(F, X):
        ref = orc.bin_level(level, X)
        lst = ref["patches"][0]
        f = [np.zeros(pg.side_shape(a)) for a in range(2)]
        orc.side_spread(kernel, pg, f, X, F, lst["all_idx"], lst["all_shift"])
        return f

    u_n = u_field(1.0)
    interior = [tuple(slice(g, sh - g) for sh in pg.side_shape(a)) for a in range(2)]
    # ---- oracle
    U_o = oracle_interp(u_n, X0)
    Xs_o = X0 + dt * U_o
    Xh_o = 0.5 * X0 + 0.5 * Xs_o
    F_o = orc.lagrangian_force(2, Xh_o, U_o, springs=(m, s, k, r))
    f_o = oracle_spread(F_o, Xh_o)
    # ---- device (marker data resident; only u and f cross the seam, as with a CPU fluid solver)
    for a in range(2):
        ib.grid_upload("u", 0, a, u_n[a])
    ib.beginDataRedistribution()
    ib.preprocessIntegrateData()
    ib.interpolateVelocity(fill_halo=True)
    ib.forwardEulerStep(dt)
    ib.beginDataRedistribution()
    ib.computeLagrangianForce()
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=True)
    f_d = [ib.grid_download("f", 0, a) for a in range(2)]
    for a in range(2):
        scale = np.max(np.abs(f_o[a][interior[a]]))
        assert np.max(np.abs(f_d[a][interior[a]] - f_o[a][interior[a]])) <= 1e-12 * scale
    # stand-in fluid update from the ORACLE's f on both sides, so the second half compares like with like
    c = 1e-4
    u_h = [u_n[a] + c * f_o[a] for a in range(2)]
    for a in range(2):  # periodic wrap of the host field's ghost cells is the device's job: upload interiors, fill halo
        ib.grid_upload("u", 0, a, u_h[a])
    ib.interpolateVelocity(fill_halo=True)
    ib.midpointStep(dt)
    ib.postprocessIntegrateData()
    # oracle second half: ghost cells of u_h must hold the periodic images (what fill_halo does on the device)
    u_h_filled = []
    for a in range(2):
        arr = u_h[a].copy()
        n0 = [n + (1 if d == a else 0) for d in range(2)]
        idx = [np.mod(np.arange(arr.shape[1 - d]) - g, n) + g for d in range(2)]
        u_h_filled.append(np.ascontiguousarray(arr[np.ix_(idx[1], idx[0])]))
    U2_o = oracle_interp(u_h_filled, Xh_o)
    Xn_o = X0 + dt * U2_o
    assert np.max(np.abs(ib.getLData("F") - F_o)) <= 1e-12 * np.max(np.abs(F_o))
    assert np.max(np.abs(ib.getLData("U") - U2_o)) <= 1e-12 * np.max(np.abs(U2_o))
    assert np.max(np.abs(ib.getLData("X") - Xn_o)) <= 1e-14
    ib.close()
