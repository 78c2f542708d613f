"""BASELINE.json's configurations as GPU parity cases (reduced sizes where the oracle would take minutes):
C1  examples/IB/explicit/ex1: 2D elastic ellipse (curve2d_64.vertex), IB_4, 64^2 periodic grid, one patch
C2  3D spherical shell (Fibonacci lattice), IB_4, one periodic patch: many markers per cell
C3  IB_6, uniform + clustered shell with radial jitter, 2x2x1 patches
C4  two-level AMR: markers on the finest level whose patches do not cover the domain; ghost cells with no
    same-level owner are dropped
Each compares spreadForce / interpolateVelocity on the resident level with the oracle's model of the reference
path (redundant ghost-box spreading with periodic shifts, interiors kept; interpolation at interior lists)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit


def _uniform(seed, n, lo, hi):
    return lo + (hi - lo) * splitmix64_unit(seed, np.arange(n))

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def api():
    from ibamr_b200 import api as _api
    _api.default_context()
    return _api


def relerr(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def _side_field(pg, ncell, x_lower, L, seed, ndim):
    out = []
    for axis in range(ndim):
        c = pg.side_coords(axis)
        f = np.sin(2 * np.pi * (c[axis] - x_lower[axis]) / L[axis]) * np.cos(2 * np.pi * (c[(axis + 1) % ndim] - x_lower[(axis + 1) % ndim]) / L[(axis + 1) % ndim])
        gi = []
        for d in range(ndim):
            cnt = pg.upper[d] - pg.lower[d] + 1 + (1 if d == axis else 0) + 2 * pg.gcw[d]
            gi.append(np.mod(np.arange(cnt) + pg.lower[d] - pg.gcw[d], ncell[d]))
        mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
        lin = np.zeros(f.shape, dtype=np.int64)
        mul = 1
        for d in range(ndim):
            lin += mesh[d] * mul
            mul *= ncell[d]
        out.append(np.ascontiguousarray(f + 1e-3 * splitmix64_unit(seed + axis, lin.reshape(-1)).reshape(f.shape)))
    return out


def _run_level_case(api, level, kernel, X, F, periodic_fill=True, check_ghost_fill=True):
    """Runs the resident level and the reference model; returns the max relative errors."""
    ndim = level.ndim
    g = level.gcw[0]
    N = X.shape[0]
    L = tuple(level.x_upper[d] - level.x_lower[d] for d in range(ndim))
    Xw, _ = orc.wrap_positions(X, level.x_lower, level.x_upper, level.periodic)
    Xw = Xw.reshape(-1, ndim)
    ref = orc.bin_level(level, Xw)
    ib = api.IBMethodB200(ndim, level.domain_lower, level.domain_upper(), level.x_lower, level.x_upper, level.periodic, level.boxes,
                          gcw=g, kernel_fcn=kernel)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    cells, owner = ib.getCellsAndOwners()
    assert np.array_equal(cells, ref["cells"]) and np.array_equal(owner, ref["owner"])
    U_ref = np.zeros((N, ndim))
    f_ref = []
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        u = _side_field(pg, level.domain_ncells, level.x_lower, L, 500, ndim)
        lst = ref["patches"][p]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, ndim)[lst["interior_mask"]]
        orc.side_interp(kernel, pg, u, Xw, ii, sh.reshape(-1), U_ref)
        fo = [np.zeros(pg.side_shape(a)) for a in range(ndim)]
        orc.side_spread(kernel, pg, fo, Xw, F, lst["all_idx"], lst["all_shift"])
        f_ref.append(fo)
        for a in range(ndim):
            ib.grid_upload("u", p, a, u[a])  # analytic values everywhere (ghosts included)
    ib.interpolateVelocity(fill_halo=periodic_fill)
    U = ib.getLData("U")
    owned = ref["owner"] >= 0
    eu = relerr(U[owned], U_ref[owned])
    ib.spreadForce(accumulate_halo=True)
    ef = 0.0
    for p in range(len(level.boxes)):
        for a in range(ndim):
            got = ib.grid_download("f", p, a)
            sl = tuple(slice(g, s - g) for s in got.shape)
            ef = max(ef, np.max(np.abs(got[sl] - f_ref[p][a][sl])) / max(max(np.max(np.abs(x)) for x in f_ref[p]), 1e-300))
    ib.close()
    return eu, ef


def test_c1_ex1_ellipse_2d(api, golden_dir):
    """C1: the 304 vertices of examples/IB/explicit/ex1/curve2d_64.vertex on the 64^2 periodic unit square."""
    with open(os.path.join(golden_dir, "curve2d_64.vertex")) as f:
        n = int(f.readline().split()[0])
        X = np.array([[float(t) for t in f.readline().split()[:2]] for _ in range(n)])
    assert X.shape == (304, 2)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(n)) - 1 for d in range(2)], axis=1)
    level = orc.Level(2, (0, 0), (64, 64), (0.0, 0.0), (1.0, 1.0), (1, 1), [((0, 0), (63, 63))], (3, 3))
    eu, ef = _run_level_case(api, level, "IB_4", X, F)
    assert eu <= TOL and ef <= TOL


def test_c2_spherical_shell_dense(api):
    """C2 (reduced): Fibonacci-lattice shell R = 0.25 about the centre, 60k markers on 64^3 (tens of markers per cell)."""
    N, n = 60000, 64
    k = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * k / N)
    th = np.pi * (1 + 5 ** 0.5) * k
    X = 0.5 + 0.25 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0,) * 3, (n - 1,) * 3)], (3,) * 3)
    eu, ef = _run_level_case(api, level, "IB_4", X, F)
    assert eu <= TOL and ef <= TOL


def test_c2_spherical_shell_full_size(api):
    """C2 at BASELINE.json's stated size: 2^20 markers on a shell of radius 0.25 on the 256^3 periodic unit cube (IB_4), value by
    value against the oracle (the configs were only checked at reduced size)."""
    N, n = 1 << 20, 256
    k = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * k / N)
    th = np.pi * (1 + 5 ** 0.5) * k
    X = 0.5 + 0.25 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0,) * 3, (n - 1,) * 3)], (3,) * 3)
    eu, ef = _run_level_case(api, level, "IB_4", X, F)
    assert eu <= TOL and ef <= TOL


def test_c3_ib6_uniform_plus_jittered_shell(api):
    """C3 (reduced): IB_6, uniform markers + shell R = 0.3 with radial jitter N(0, h), 2x2x1 patches of a 48^3 grid."""
    n, N = 48, 30000
    h = 1.0 / n
    Xu = np.stack([splitmix64_unit(3 + 10 * d, np.arange(N // 2)) for d in range(3)], axis=1)
    u1, u2 = splitmix64_unit(4, np.arange(N // 2)), splitmix64_unit(5, np.arange(N // 2))
    jitter = np.sqrt(-2 * np.log(np.maximum(u1, 1e-300))) * np.cos(2 * np.pi * u2) * h  # Box-Muller
    k = np.arange(N // 2) + 0.5
    phi = np.arccos(1 - 2 * k / (N // 2))
    th = np.pi * (1 + 5 ** 0.5) * k
    Xs = 0.5 + (0.3 + jitter)[:, None] * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    X = np.concatenate([Xu, Xs])
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    boxes = [((0, 0, 0), (23, 23, 47)), ((24, 0, 0), (47, 23, 47)), ((0, 24, 0), (23, 47, 47)), ((24, 24, 0), (47, 47, 47))]
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (4,) * 3)
    eu, ef = _run_level_case(api, level, "IB_6", X, F)
    assert eu <= TOL and ef <= TOL


def test_c4_two_level_amr_fine_patches(api):
    """C4 (reduced): the finest level (ratio 4 of a 16^3 coarse grid -> 64^3 index space) exists only over the
    central region, 2x2x2 patches of 16^3; markers on a sphere at least gcw fine cells inside it
    (IBMethod::setupTagBuffer, IBMethod.cpp:272-293).  Ghost cells facing the coarse level have no same-level
    owner: what is spread there is dropped, and u there is whatever the CF interpolation put (here: analytic)."""
    nfine, N = 64, 20000
    boxes = []
    for kz in range(2):
        for ky in range(2):
            for kx in range(2):
                lo = (16 + 16 * kx, 16 + 16 * ky, 16 + 16 * kz)
                boxes.append((lo, tuple(l + 15 for l in lo)))
    level = orc.Level(3, (0,) * 3, (nfine,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (3,) * 3)
    k = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * k / N)
    th = np.pi * (1 + 5 ** 0.5) * k
    X = 0.5 + 0.18 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    # fill_halo=True copies same-level neighbours into ghosts; CF-facing ghosts keep the uploaded analytic values
    eu, ef = _run_level_case(api, level, "IB_4", X, F)
    assert eu <= TOL and ef <= TOL


def test_c4_two_level_amr_full_size(api):
    """C4 at BASELINE.json's stated size: coarse 128^3, ratio 4 (512^3 effective), the fine level = 8 patches of 128^3 over
    the central 256^3 fine cells, 2^20 markers on a sphere inside it; value by value against the oracle."""
    nfine, N, w = 512, 1 << 20, 128
    boxes = []
    for kz in range(2):
        for ky in range(2):
            for kx in range(2):
                lo = (128 + w * kx, 128 + w * ky, 128 + w * kz)
                boxes.append((lo, tuple(l + w - 1 for l in lo)))
    level = orc.Level(3, (0,) * 3, (nfine,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (3,) * 3)
    k = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * k / N)
    th = np.pi * (1 + 5 ** 0.5) * k
    X = 0.5 + 0.2 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    eu, ef = _run_level_case(api, level, "IB_4", X, F)
    assert eu <= TOL and ef <= TOL


@pytest.mark.parametrize("kernel", ["BSPLINE_4", "PIECEWISE_LINEAR"])
def test_nonperiodic_walls(api, kernel):
    """Wall-bounded domain: markers close to the walls spread into ghost cells outside the domain, which have no
    owner and, with no boundary conditions registered (ibk_level_set_wall_bc), are dropped; interiors must still equal the
    reference's plain spread.  (The fold-back itself: test_walls_fold_back_the_spread_force.)"""
    n, N = 24, 8000
    g = orc.min_ghost_width(kernel)
    boxes = [((0, 0, 0), (11, 23, 23)), ((12, 0, 0), (23, 23, 23))]
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (0, 0, 0), boxes, (g,) * 3)
    X = np.stack([0.01 + 0.98 * splitmix64_unit(20 + d, np.arange(N)) for d in range(3)], axis=1)
    F = np.stack([2 * splitmix64_unit(1 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    eu, ef = _run_level_case(api, level, kernel, X, F)
    assert eu <= TOL and ef <= TOL


@pytest.mark.parametrize("kernel,bc", [("IB_4", "dirichlet"), ("IB_4", "robin"), ("BSPLINE_3", "dirichlet"), ("IB_6", "mixed")])
def test_walls_fold_back_the_spread_force(api, kernel, bc):
    """N3, wall part: a channel (periodic in x and y, walls in z) cut into two patches.  With boundary conditions registered
    the force spread into the ghost cells outside the domain is folded back by the adjoint of the Robin extrapolation
    (CartSideRobinPhysBdryOp::accumulateFromPhysicalBoundaryData) instead of being dropped.  Reference model: every patch
    spreads all markers of its ghost box (periodic images included) into a zeroed array, then folds its own walls
    (LDataManager.cpp:623-657); interiors are compared."""
    n, N = 32, 12000
    g = orc.min_ghost_width(kernel)
    boxes = [((0, 0, 0), (15, 31, 31)), ((16, 0, 0), (31, 31, 31))]
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 0), boxes, (g,) * 3)
    i = np.arange(N)
    X = np.stack([splitmix64_unit(20, i), splitmix64_unit(21, i), 0.005 + 0.99 * splitmix64_unit(22, i)], axis=1)
    X[: N // 2, 2] = np.where(splitmix64_unit(23, i[: N // 2]) < 0.5, 0.002 + 0.08 * splitmix64_unit(24, i[: N // 2]),
                              0.998 - 0.08 * splitmix64_unit(25, i[: N // 2]))  # half of them within a few cells of a wall
    F = np.stack([2 * splitmix64_unit(1 + d, i) - 1 for d in range(3)], axis=1)
    a = np.ones((3, 2, 3))
    b = np.zeros((3, 2, 3))
    if bc == "robin":
        b[:] = 0.5
    elif bc == "mixed":
        b[2, 0, :] = 0.25          # lower wall Robin, upper wall Dirichlet
        a[2, 0, :] = 2.0
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 0), level.boxes, gcw=g, kernel_fcn=kernel)
    ib.setWallBc(a, b)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=True)
    ref = orc.bin_level(level, X)
    worst = 0.0
    for p in range(2):
        pg = level.patch_geom(p)
        lst = ref["patches"][p]
        fr = [np.zeros(pg.side_shape(c)) for c in range(3)]
        orc.side_spread(kernel, pg, fr, X, F, lst["all_idx"], lst["all_shift"])
        before = [x.copy() for x in fr]
        orc.fold_walls(level, p, fr, a, b)
        assert any(np.max(np.abs(fr[c] - before[c])) > 1e-6 for c in range(3)), "the case must exercise the fold-back"
        for c in range(3):
            got = ib.grid_download("f", p, c)
            sl = tuple(slice(g, s - g) for s in got.shape)
            worst = max(worst, np.max(np.abs(got[sl] - fr[c][sl])) / np.max(np.abs(fr[c][sl])))
    ib.close()
    assert worst <= TOL
    # walls in two dimensions need the co-dimension two extrapolation: refused, not silently wrong
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 0, 0), level.boxes, gcw=g, kernel_fcn=kernel)
    with pytest.raises(api.IBKError):
        ib.setWallBc(a, b)
    ib.close()


@pytest.mark.parametrize("kernel", ["IB_4", "IB_6", "BSPLINE_3", "BSPLINE_4", "PIECEWISE_LINEAR", "IB_3", "BSPLINE_5", "BSPLINE_6",
                                    "PIECEWISE_CUBIC", "IB_5", "PIECEWISE_CONSTANT", "COMPOSITE_BSPLINE_32", "COMPOSITE_BSPLINE_43",
                                    "COMPOSITE_BSPLINE_45", "COMPOSITE_BSPLINE_56", "DISCONTINUOUS_LINEAR", "IB_4_W8"])
def test_dense_bricks_every_kernel(api, kernel):
    """A structure-like cloud: 40k markers in a slab a few cells thick that crosses patch boundaries and the periodic
    boundary, i.e. hundreds of markers per 4^3-cell brick: these bricks take spread_dense_kernel (register
    footprints of 6^3 / 8^3 / 10^3 points), their sparse neighbours the tile kernel; together they must equal the
    reference's serial spreading."""
    n, N = 32, 40000
    g = orc.min_ghost_width(kernel)
    i = np.arange(N)
    X = np.stack([splitmix64_unit(41, i), splitmix64_unit(42, i), 0.97 + 0.08 * splitmix64_unit(43, i)], axis=1)  # z in [0.97, 1.05): wraps
    X[: N // 4, 0] = 0.48 + 0.05 * splitmix64_unit(44, i[: N // 4])  # and a denser streak across the patch boundary at x = 0.5
    F = np.stack([2 * splitmix64_unit(51 + d, i) - 1 for d in range(3)], axis=1)
    boxes = [((0, 0, 0), (15, 31, 31)), ((16, 0, 0), (31, 31, 31))]
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (g,) * 3)
    eu, ef = _run_level_case(api, level, kernel, X, F)
    assert eu <= TOL and ef <= TOL


@pytest.mark.parametrize("kernel", ["IB_4", "IB_6"])
def test_many_stencils_outside_their_tile_are_still_spread_exactly(api, kernel):
    """ADVICE r1 (silent overflow of the 4096-entry exception list): positions that moved since the binning put thousands of
    stencils outside the accumulator of their tile.  They must all be spread (by the fix-up, in sorted order), none dropped.
    The positions are changed behind the library's back through the device pointer of the X column (ibk_markers_upload would
    demand a re-bin), 20000 markers by 3 cells each: every one of them misses its brick's footprint."""
    import torch

    ndim, n, N = 3, 64, 30000
    g = orc.min_ghost_width(kernel)
    level = orc.Level(ndim, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0,) * 3, (n - 1,) * 3)], (g,) * 3)
    h = 1.0 / n
    X = np.stack([_uniform(901 + d, N, 8 * h, 1.0 - 8 * h) for d in range(3)], axis=1)
    F = np.stack([_uniform(911 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), level.boxes, kernel_fcn=kernel)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    order = ib.getSortedLagrangianIndices()          # storage position -> Lagrangian index
    moved = np.zeros(N, dtype=bool)
    moved[:20000] = True
    Xn = X.copy()
    Xn[moved] += np.array([3 * h, -3 * h, 3 * h])
    ptr, stride = ib.marker_device_ptr("X")
    Xs = np.ascontiguousarray(Xn[order].T)           # SoA in storage order
    # (raw copy into the library's column: wrap it as a torch tensor through the CUDA array interface)
    class _Dev:
        def __init__(self, p, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (p, False), "version": 2}
    ib.ctx.synchronize()
    for d in range(3):
        torch.as_tensor(_Dev(ptr + 8 * stride * d, N), device="cuda").copy_(torch.from_numpy(Xs[d]))
    torch.cuda.synchronize()
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=False)
    pg = level.patch_geom(0)
    f_ref = [np.zeros(pg.side_shape(a)) for a in range(3)]
    orc.side_spread(kernel, pg, f_ref, Xn, F, np.arange(N, dtype=np.int32), np.zeros(3 * N))
    for a in range(3):
        f = ib.grid_download("f", 0, a)
        assert np.max(np.abs(f - f_ref[a])) <= 1e-12 * np.max(np.abs(f_ref[a]))
    # and the flags are clean again: a second spread adds exactly the same once more
    ib.spreadForce(accumulate_halo=False)
    for a in range(3):
        f = ib.grid_download("f", 0, a)
        assert np.max(np.abs(f - 2.0 * f_ref[a])) <= 1e-12 * np.max(np.abs(f_ref[a])) * 2
    ib.close()
