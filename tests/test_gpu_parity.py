"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, and against the reference's golden files.  Tolerance for fields is north_star's:
max|a-b| / max|b| <= 1e-12 in fp64; binning is bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import (read_ghost_accumulation_golden, read_interpolate_golden, splitmix64_unit, std_uniform_stream)

pytestmark = pytest.mark.gpu

KERNELS = ["IB_4", "IB_6", "BSPLINE_3", "BSPLINE_4", "PIECEWISE_LINEAR", "IB_3", "BSPLINE_5", "BSPLINE_6", "PIECEWISE_CUBIC", "IB_5",
           "PIECEWISE_CONSTANT", "COMPOSITE_BSPLINE_32", "COMPOSITE_BSPLINE_23", "COMPOSITE_BSPLINE_43", "COMPOSITE_BSPLINE_34",
           "COMPOSITE_BSPLINE_54", "COMPOSITE_BSPLINE_45", "COMPOSITE_BSPLINE_65", "COMPOSITE_BSPLINE_56", "DISCONTINUOUS_LINEAR",
           "IB_4_W8"]
TOL = 1e-12


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def api():
    from ibamr_b200 import api as _api
    _api.default_context()  # raises loudly if libibk.so or the GPU is missing
    return _api


def _uniform(seed, n, lo, hi):
    return lo + (hi - lo) * splitmix64_unit(seed, np.arange(n))


# ------------------------------------------------------------------------------------------------
# seam B4: raw funnel vs the Fortran restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", KERNELS)
def test_raw_interp_spread_vs_oracle(api, kernel, ndim):
    """One array, depth 2, markers everywhere incl. the ghost region and beyond (exercises the
    clipping to the ghost box), a shuffled index list with duplicates removed and periodic shifts."""
    g = orc.min_ghost_width(kernel)
    ilower = (3, -2, 5)[:ndim]
    iupper = (3 + 20, -2 + 17, 5 + 13)[:ndim]
    nugc = (g, g + 1, g)[:ndim]
    dx = (0.05, 0.04, 0.0625)[:ndim]
    x_lower = (-0.3, 0.1, 0.25)[:ndim]
    x_upper = tuple(x_lower[d] + dx[d] * (iupper[d] - ilower[d] + 1) for d in range(ndim))
    depth = 2
    shape = (depth,) + tuple(reversed([iupper[d] - ilower[d] + 1 + 2 * nugc[d] for d in range(ndim)]))
    rng = np.random.default_rng(1234 + ndim)
    u = rng.standard_normal(shape)
    n = 4000
    X = np.stack([_uniform(7 + d, n, x_lower[d] - (nugc[d] + 1.5) * dx[d], x_upper[d] + (nugc[d] + 1.5) * dx[d])
                  for d in range(ndim)], axis=1)
    indices = rng.permutation(n)[:3000].astype(np.int32)
    Xshift = np.zeros((indices.size, ndim))
    Xshift[::7, 0] = dx[0] * 3  # some entries carry a (periodic-image) shift
    Xshift[::11, ndim - 1] = -dx[ndim - 1] * 2
    # interpolate
    V_ref = np.full((n, depth), -7.0)
    orc.interp_raw(kernel, ndim, dx, x_lower, depth, ilower, iupper, nugc, u, indices, Xshift, X, V_ref)
    V = np.full((n, depth), -7.0)
    api.raw_interp_host(kernel, ndim, dx, x_lower, x_upper, depth, ilower, iupper, nugc, u, indices, Xshift, X, V)
    assert relerr(V, V_ref) <= TOL
    untouched = np.setdiff1d(np.arange(n), indices)
    assert np.all(V[untouched] == -7.0)  # unlisted markers are left alone, as in the reference
    # spread (u += ...)
    F = rng.standard_normal((n, depth))
    u_ref = u.copy()
    orc.spread_raw(kernel, ndim, dx, x_lower, depth, indices, Xshift, X, F, ilower, iupper, nugc, u_ref)
    u_gpu = u.copy()
    api.raw_spread_host(kernel, ndim, dx, x_lower, x_upper, depth, indices, Xshift, X, F, ilower, iupper, nugc, u_gpu)
    assert relerr(u_gpu - u, u_ref - u) <= TOL
    assert relerr(u_gpu, u_ref) <= TOL


def test_raw_empty_and_single(api):
    """Edge cases: empty index list is a no-op; a single marker; every marker outside the array."""
    ndim, kernel = 3, "IB_4"
    il, iu, ng = (0, 0, 0), (7, 7, 7), (3, 3, 3)
    dx, xl, xu = (0.125,) * 3, (0.0,) * 3, (1.0,) * 3
    u = np.arange(14 ** 3, dtype=np.float64).reshape(1, 14, 14, 14)
    X = np.array([[0.5, 0.5, 0.5], [9.0, 9.0, 9.0]])
    V = np.full((2, 1), 5.0)
    api.raw_interp_host(kernel, ndim, dx, xl, xu, 1, il, iu, ng, u, np.zeros(0, np.int32), None, X, V)
    assert np.all(V == 5.0)
    Vr = V.copy()
    for idx in ([0], [1], [0, 1]):
        idx = np.array(idx, np.int32)
        api.raw_interp_host(kernel, ndim, dx, xl, xu, 1, il, iu, ng, u, idx, None, X, V)
        orc.interp_raw(kernel, ndim, dx, xl, 1, il, iu, ng, u, idx, np.zeros(idx.size * 3), X, Vr)
        assert relerr(V, Vr) <= TOL
    assert V[1, 0] == 0.0  # far outside: the whole stencil is clipped, the reference writes 0


# ------------------------------------------------------------------------------------------------
# seam B3 against the reference's golden files
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_interpolate_01_3d(api, kernel, golden_dir):
    """tests/interpolate/interpolate_01.cpp (3D, multilinear field) through LEInteractor.interpolate."""
    gold = read_interpolate_golden(os.path.join(golden_dir, f"interpolate_01_3d.{kernel.lower()}.output"))
    N = 8
    g = api.LEInteractor.getMinimumGhostWidth(kernel)
    box = api.Box((N,) * 3, (2 * N - 1,) * 3)
    patch = api.Patch(box, (0.25,) * 3, (0.5,) * 3, (1.0 / (4 * N),) * 3)
    pg = orc.PatchGeom(box.lower, box.upper, patch.x_lower, patch.x_upper, patch.dx, (g,) * 3)
    x, y, z = pg.cell_coords()
    f = 1 + 2 * x + 3 * y - z + 4 * x * y + 2 * x * z + 3 * x * y * z
    q = api.CellData(box, 3, g, np.stack([f, f, f]))
    X = std_uniform_stream(42, 300, 0.25, 0.5).reshape(100, 3)
    Q = np.full((100, 3), np.finfo(np.float64).max)
    api.LEInteractor.interpolate(Q, 3, X, 3, q, patch, box, kernel)
    # (BSPLINE_6's degree-5 polynomial amplifies evaluation-order rounding, see tests/test_oracle_golden.py)
    np.testing.assert_allclose(Q, gold[:, 3:6], rtol=0,
                               atol=1e-11 if kernel in ("BSPLINE_6", "COMPOSITE_BSPLINE_65", "COMPOSITE_BSPLINE_56") else 2e-12)


@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_interpolate_01_2d(api, kernel, golden_dir):
    gold = read_interpolate_golden(os.path.join(golden_dir, f"interpolate_01_2d.{kernel.lower()}.output"))
    N = 16
    g = api.LEInteractor.getMinimumGhostWidth(kernel)
    box = api.Box((N,) * 2, (2 * N - 1,) * 2)
    patch = api.Patch(box, (0.25,) * 2, (0.5,) * 2, (1.0 / (4 * N),) * 2)
    pg = orc.PatchGeom(box.lower, box.upper, patch.x_lower, patch.x_upper, patch.dx, (g,) * 2)
    x, y = pg.cell_coords()
    f = np.sin(2 * np.pi * (x - 0.1234)) * np.sin(2 * np.pi * (y - 0.1234))
    q = api.CellData(box, 2, g, np.stack([f, f]))
    X = std_uniform_stream(42, 200, 0.25, 0.5).reshape(100, 2)
    Q = np.full((100, 2), np.finfo(np.float64).max)
    api.LEInteractor.interpolate(Q, 2, X, 2, q, patch, box, kernel)
    reach = {"IB_4": 2, "IB_6": 3, "BSPLINE_3": 2, "BSPLINE_4": 2, "PIECEWISE_LINEAR": 1, "IB_3": 2, "BSPLINE_5": 3, "BSPLINE_6": 3,
             "PIECEWISE_CUBIC": 2, "IB_5": 3, "PIECEWISE_CONSTANT": 1, "COMPOSITE_BSPLINE_32": 2, "COMPOSITE_BSPLINE_23": 2,
             "COMPOSITE_BSPLINE_43": 2, "COMPOSITE_BSPLINE_34": 2, "COMPOSITE_BSPLINE_54": 3, "COMPOSITE_BSPLINE_45": 3,
             "COMPOSITE_BSPLINE_65": 3, "COMPOSITE_BSPLINE_56": 3, "DISCONTINUOUS_LINEAR": 1, "IB_4_W8": 4}[kernel]
    cell = np.floor((X - 0.25) / patch.dx[0]).astype(int)
    inside = np.all((cell - reach >= 0) & (cell + reach <= N - 1), axis=1)
    np.testing.assert_allclose(Q[inside], gold[inside][:, 3:5], rtol=0, atol={"BSPLINE_6": 2e-12, "BSPLINE_5": 2e-13, "COMPOSITE_BSPLINE_65": 2e-12, "COMPOSITE_BSPLINE_56": 2e-12,
                                     "COMPOSITE_BSPLINE_54": 2e-13, "COMPOSITE_BSPLINE_45": 2e-13}.get(kernel, 5e-14))
    # and everywhere (ghost data included) against the oracle on identical inputs
    Qo = orc.cell_interp_positions(kernel, pg, q.array, 2, X)
    assert relerr(Q, Qo) <= TOL


GA_SIDE = {"2d.side.spread.a": (2, 8, True), "2d.side.spread.b": (2, 8, True), "3d.side.spread.a": (3, 8, False),
           "3d.side.spread.b": (3, 8, False)}


@pytest.mark.parametrize("case", sorted(GA_SIDE))
def test_golden_ghost_accumulation_resident_level(api, case, golden_dir):
    """tests/IBTK/ghost_accumulation_01.cpp (side, spread): owner-only PIECEWISE_LINEAR spreading into
    interior + ghost cells of every patch followed by the ghost accumulation, through the resident
    level (IBMethodB200.spreadForce with the device halo sum)."""
    ndim, n, periodic = GA_SIDE[case]
    gold = read_ghost_accumulation_golden(os.path.join(golden_dir, f"ghost_accumulation_01_{case}.output"), ndim)
    dx = 1.0 / n
    boxes = []
    for p in gold:
        lo = tuple(int(round(p["x_lower"][d] / dx)) for d in range(ndim))
        hi = tuple(int(round(p["x_upper"][d] / dx)) - 1 for d in range(ndim))
        boxes.append((lo, hi))
    stream = std_uniform_stream(42, 2 * ndim * 100)
    Q = stream[:ndim * 100].reshape(100, ndim)
    X = stream[ndim * 100:].reshape(100, ndim)
    ib = api.IBMethodB200(ndim, (0,) * ndim, (n - 1,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (periodic,) * ndim, boxes, gcw=2,
                          kernel_fcn="PIECEWISE_LINEAR")
    ib.setPositions(X)
    ib.setLData("F", Q)
    ib.beginDataRedistribution()
    ib.endDataRedistribution()
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=True)
    checked = 0
    for p, gp in enumerate(gold):
        lo, hi = boxes[p]
        for axis, entries in gp["comps"].items():
            a = ib.grid_download("f", p, axis)
            for idx, val in entries.items():
                loc = tuple(idx[d] - lo[d] + 2 for d in reversed(range(ndim)))
                assert abs(a[loc] - val) <= 1e-12 * max(1.0, abs(val)), (case, p, axis, idx, a[loc], val)
                checked += 1
    assert checked > 100
    ib.close()


@pytest.mark.parametrize("case", ["2d.cell.spread.a", "3d.cell.spread.a", "3d.cell.spread.b"])
def test_golden_ghost_accumulation_cell(api, case, golden_dir):
    """CellData variants: LEInteractor.spread on the GPU, accumulation by the oracle's restatement."""
    ndim = int(case[0])
    periodic = ndim == 2
    n = 8
    gold = read_ghost_accumulation_golden(os.path.join(golden_dir, f"ghost_accumulation_01_{case}.output"), ndim)
    dx = 1.0 / n
    boxes = []
    for p in gold:
        lo = tuple(int(round(p["x_lower"][d] / dx)) for d in range(ndim))
        hi = tuple(int(round(p["x_upper"][d] / dx)) - 1 for d in range(ndim))
        boxes.append((lo, hi))
    level = orc.Level(ndim, (0,) * ndim, (n,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (int(periodic),) * ndim, boxes,
                      (2,) * ndim)
    stream = std_uniform_stream(42, (1 + ndim) * 100)
    Q = stream[:100].reshape(100, 1)
    X = stream[100:].reshape(100, ndim)
    arrays = []
    for p, (lo, hi) in enumerate(boxes):
        pg = level.patch_geom(p)
        box = api.Box(lo, hi)
        patch = api.Patch(box, pg.x_lower, pg.x_upper, pg.dx, touches_regular_bdry=not periodic)
        q = api.CellData(box, 1, 2)
        api.LEInteractor.spread(q, Q, 1, X, ndim, patch, box, "PIECEWISE_LINEAR")
        arrays.append([q.array[0]])
    orc.ghost_accumulate(level, arrays, "cell")
    checked = 0
    for p, gp in enumerate(gold):
        lo, hi = boxes[p]
        for axis, entries in gp["comps"].items():
            for idx, val in entries.items():
                loc = tuple(idx[d] - lo[d] + 2 for d in reversed(range(ndim)))
                assert abs(arrays[p][0][loc] - val) <= 1e-12 * max(1.0, abs(val))
                checked += 1
    assert checked > 0


# ------------------------------------------------------------------------------------------------
# seam B3 vs oracle: SideData, position-only with a box, index-set form with periodic shifts
# ------------------------------------------------------------------------------------------------
def _side_fields(pg, seed):
    out = []
    for axis in range(pg.ndim):
        c = pg.side_coords(axis)
        f = np.sin(2 * np.pi * c[axis]) * np.cos(2 * np.pi * c[(axis + 1) % pg.ndim])
        f = f + 1e-3 * splitmix64_unit(seed + axis, np.arange(f.size)).reshape(f.shape)
        out.append(np.ascontiguousarray(f))
    return out


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", ["IB_4", "IB_6", "BSPLINE_3", "PIECEWISE_LINEAR", "BSPLINE_5"])
def test_node_and_edge_centerings_vs_oracle(api, kernel, ndim):
    """NodeData (depth 2, shifted in every dimension) and EdgeData (per axis, shifted in every dimension but the axis)
    position-only interpolate / spread (LEInteractor.cpp:2983-3043, 3260-3340, 4122-4186, 4386-4466) against the oracle
    driven with the same shifted array geometry; markers also outside the patch (not listed: untouched)."""
    g = orc.min_ghost_width(kernel)
    n = 20
    lo, hi = (4,) * ndim, (4 + n - 1,) * ndim
    dx = (1.0 / 32,) * ndim
    xl = tuple(dx[d] * lo[d] for d in range(ndim))
    xu = tuple(dx[d] * (hi[d] + 1) for d in range(ndim))
    pg = orc.PatchGeom(lo, hi, xl, xu, dx, (g,) * ndim)
    box = api.Box(lo, hi)
    patch = api.Patch(box, xl, xu, dx)
    N = 4000
    X = np.stack([_uniform(53 + d, N, xl[d] - 2 * dx[d], xu[d] + 2 * dx[d]) for d in range(ndim)], axis=1)
    sentinel = 123.25
    # ---- node
    depth = 2
    qn = api.NodeData(box, depth, g)
    qn.array[...] = _uniform(61, qn.array.size, -1.0, 1.0).reshape(qn.array.shape)
    Qo = orc.node_interp_positions(kernel, pg, qn.array, depth, X)
    Q = np.full((N, depth), np.finfo(np.float64).max)
    api.LEInteractor.interpolate(Q, depth, X, ndim, qn, patch, box, kernel)
    assert relerr(Q, Qo) <= TOL
    F = np.stack([_uniform(70 + d, N, -1.0, 1.0) for d in range(depth)], axis=1)
    fo = np.zeros_like(qn.array)
    orc.node_spread_positions(kernel, pg, fo, depth, X, F)
    qs = api.NodeData(box, depth, g)
    api.LEInteractor.spread(qs, F, depth, X, ndim, patch, box, kernel)
    assert relerr(qs.array, fo) <= TOL and np.any(fo != 0.0)
    # ---- edge
    qe = api.EdgeData(box, 1, g)
    for a in range(ndim):
        qe.arrays[a][...] = _uniform(80 + a, qe.arrays[a].size, -1.0, 1.0).reshape(qe.arrays[a].shape)
    Qo = orc.edge_interp_positions(kernel, pg, qe.arrays, X)
    Q = np.full((N, ndim), np.finfo(np.float64).max)
    api.LEInteractor.interpolate(Q, ndim, X, ndim, qe, patch, box, kernel)
    assert relerr(Q, Qo) <= TOL
    Fe = np.stack([_uniform(90 + d, N, -1.0, 1.0) for d in range(ndim)], axis=1)
    eo = [np.zeros_like(a) for a in qe.arrays]
    orc.edge_spread_positions(kernel, pg, eo, X, Fe)
    qs = api.EdgeData(box, 1, g)
    api.LEInteractor.spread(qs, Fe, ndim, X, ndim, patch, box, kernel)
    for a in range(ndim):
        assert relerr(qs.arrays[a], eo[a]) <= TOL and np.any(eo[a] != 0.0)
    # edge data is vector-valued: a scalar Q is refused like the reference does
    with pytest.raises(api.IBKError) as e:
        api.LEInteractor.interpolate(np.zeros((N, 1)), 1, X, ndim, qe, patch, box, kernel)
    assert e.value.code == api.IBK_ERR_DEPTH


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kernel", KERNELS)
def test_side_positions_vs_oracle(api, kernel, ndim):
    g = orc.min_ghost_width(kernel)
    n = 24
    lo, hi = (8,) * ndim, (8 + n - 1,) * ndim
    dx = (1.0 / 64,) * ndim
    xl = tuple(dx[d] * lo[d] for d in range(ndim))
    xu = tuple(dx[d] * (hi[d] + 1) for d in range(ndim))
    pg = orc.PatchGeom(lo, hi, xl, xu, dx, (g,) * ndim)
    box = api.Box(lo, hi)
    patch = api.Patch(box, xl, xu, dx)
    u = _side_fields(pg, 100)
    N = 5000
    X = np.stack([_uniform(3 + d, N, xl[d] - 2 * dx[d], xu[d] + 2 * dx[d]) for d in range(ndim)], axis=1)
    # a sub-box of the patch as interpolation box
    sub = api.Box(tuple(l + 2 for l in lo), tuple(h - 3 for h in hi))
    for b in (box, sub):
        Qo = orc.side_interp_positions(kernel, pg, u, X, b.lower, b.upper)
        q = api.SideData(box, 1, g, u)
        Q = np.zeros((N, ndim))
        api.LEInteractor.interpolate(Q, ndim, X, ndim, q, patch, b, kernel)
        assert relerr(Q, Qo) <= TOL
    F = np.stack([_uniform(20 + d, N, -1.0, 1.0) for d in range(ndim)], axis=1)
    fo = [np.zeros_like(a) for a in u]
    orc.side_spread_positions(kernel, pg, fo, X, F, box.lower, box.upper)
    q = api.SideData(box, 1, g)
    api.LEInteractor.spread(q, F, ndim, X, ndim, patch, box, kernel)
    for axis in range(ndim):
        assert relerr(q.arrays[axis], fo[axis]) <= TOL


@pytest.mark.parametrize("kernel", ["IB_4", "BSPLINE_4", "BSPLINE_6", "PIECEWISE_CUBIC"])
def test_side_indexed_with_periodic_shifts(api, kernel):
    """Index-set overloads: the reference's redundant-spreading design on a periodic level split in
    2x2x1 patches, lists and shifts from the LIndexSetData restatement."""
    ndim, n = 3, 16
    g = orc.min_ghost_width(kernel)
    boxes = [((0, 0, 0), (7, 7, 15)), ((8, 0, 0), (15, 7, 15)), ((0, 8, 0), (7, 15, 15)), ((8, 8, 0), (15, 15, 15))]
    level = orc.Level(ndim, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (g,) * 3)
    N = 3000
    X = np.stack([_uniform(31 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(41 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    bins = orc.bin_level(level, X)
    for p in range(len(boxes)):
        pg = level.patch_geom(p)
        box = api.Box(*boxes[p])
        patch = api.Patch(box, pg.x_lower, pg.x_upper, pg.dx)
        lst = bins["patches"][p]
        # spread from the ghost-box list (LDataManager.cpp:623-652)
        fo = [np.zeros(pg.side_shape(a)) for a in range(3)]
        orc.side_spread(kernel, pg, fo, X, F, lst["all_idx"], lst["all_shift"])
        q = api.SideData(box, 1, g)
        api.LEInteractor.spread_indexed(q, F, X, lst["all_idx"], lst["all_shift"], patch, kernel)
        for a in range(3):
            assert relerr(q.arrays[a], fo[a]) <= TOL
        # interpolate at the interior list (:748-802)
        u = _side_fields(pg, 200 + p)
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, 3)[lst["interior_mask"]]
        Qo = orc.side_interp(kernel, pg, u, X, ii, sh.reshape(-1))
        Q = np.zeros((N, 3))
        api.LEInteractor.interpolate_indexed(Q, X, ii, sh.reshape(-1), api.SideData(box, 1, g, u), patch, kernel)
        assert relerr(Q, Qo) <= TOL


def test_error_behaviour(api):
    """The reference's TBOX_ERROR conditions become error codes (LEInteractor.cpp:2425-2429, 4488-4498)."""
    box = api.Box((0, 0), (7, 7))
    patch = api.Patch(box, (0.0, 0.0), (1.0, 1.0), (0.125, 0.125))
    X = np.full((4, 2), 0.5)
    Q = np.zeros((4, 2))
    with pytest.raises(api.IBKError) as e:
        api.LEInteractor.interpolate(Q, 2, X, 2, api.SideData(box, 1, 1), patch, box, "IB_4")  # gcw 1 < 3
    assert e.value.code == api.IBK_ERR_GHOST_WIDTH
    with pytest.raises(api.IBKError) as e:
        api.LEInteractor.interpolate(np.zeros((4, 1)), 1, X, 2, api.SideData(box, 1, 3), patch, box, "IB_4")
    assert e.value.code == api.IBK_ERR_DEPTH
    with pytest.raises(api.IBKError) as e:
        api.LEInteractor.interpolate(Q, 2, X, 2, api.SideData(box, 1, 3), patch, box, "IB_7")  # no such kernel (LEInteractor.cpp:2038-2050)
    assert e.value.code == api.IBK_ERR_UNKNOWN_KERNEL
    api.LEInteractor.interpolate(Q, 2, X, 2, api.SideData(box, 1, 3), patch, box, "USER_DEFINED")  # the default callback: the 4-point function
    # spread with too few ghosts is only an error at a physical boundary (:5250-5266)
    api.LEInteractor.spread(api.SideData(box, 1, 1), Q, 2, X, 2, patch, box, "PIECEWISE_LINEAR")
    bpatch = api.Patch(box, (0.0, 0.0), (1.0, 1.0), (0.125, 0.125), touches_regular_bdry=True)
    with pytest.raises(api.IBKError) as e:
        api.LEInteractor.spread(api.SideData(box, 1, 1), Q, 2, X, 2, bpatch, box, "PIECEWISE_LINEAR")
    assert e.value.code == api.IBK_ERR_GHOST_WIDTH


# ------------------------------------------------------------------------------------------------
# resident level: binning (bit-exact), spread/interp vs oracle, halo, determinism, adjointness
# ------------------------------------------------------------------------------------------------
def _level_case(ndim, n, split, kernel, periodic=True):
    g = orc.min_ghost_width(kernel)
    w = [n // s for s in split]
    boxes = []
    for kz in range(split[2] if ndim == 3 else 1):
        for ky in range(split[1]):
            for kx in range(split[0]):
                k = (kx, ky, kz)[:ndim]
                boxes.append((tuple(k[d] * w[d] for d in range(ndim)), tuple((k[d] + 1) * w[d] - 1 for d in range(ndim))))
    return orc.Level(ndim, (0,) * ndim, (n,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (int(periodic),) * ndim, boxes, (g,) * ndim)


def test_binning_bit_exact(api):
    """cell = getCellIndex(X, grid_geom, ratio), owner patch, sorted order: integer results must be
    IDENTICAL to the oracle's, including adversarial positions on cell/patch/domain boundaries."""
    ndim, n = 3, 32
    level = orc.Level(3, (0,) * 3, (n,) * 3, (-1.0,) * 3, (1.0,) * 3, (1, 1, 1),
                      [((0, 0, 0), (15, 31, 31)), ((16, 0, 0), (31, 15, 31)), ((16, 16, 0), (31, 31, 31))], (3,) * 3)
    N = 20000
    X = np.stack([_uniform(51 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    # adversarial: the reference's index_utilities points, exact cell faces, +-1 ulp around them
    adv = [(-0.35355339059327373086, -0.35355339059327373086, 0.0), (-5.5511151231257827e-17, -0.49999999999999994, 0.0),
           (-0.25, -0.25, 0.0), (-2.7755575615628914e-17, -0.32322330470336308, 0.0),
           (-4.163336342344337e-17, -0.41161165235168151, 0.0)]
    faces = -1.0 + (2.0 / n) * np.arange(n)
    for f in faces[::3]:
        adv += [(f, np.nextafter(f, 2.0), np.nextafter(f, -2.0))]
    adv += [(np.nextafter(1.0, 0.0),) * 3, (-1.0,) * 3]
    X[:len(adv)] = np.array(adv)
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (-1.0,) * 3, (1.0,) * 3, (1, 1, 1), level.boxes, gcw=3)
    ib.setPositions(X)
    ib.beginDataRedistribution()
    cells, owner = ib.getCellsAndOwners()
    Xw, _ = orc.wrap_positions(X, level.x_lower, level.x_upper, level.periodic)
    ref = orc.bin_level(level, Xw)
    assert np.array_equal(cells, ref["cells"])
    assert np.array_equal(owner, ref["owner"])
    # positions after the wrap are bit-identical too
    assert np.array_equal(ib.getLData("X"), Xw.reshape(-1, 3))
    # sorted order: markers of one patch are contiguous, patches in order, within a cell by Lagrangian index
    lag = ib.getSortedLagrangianIndices()
    assert sorted(lag.tolist()) == list(range(N))
    own_sorted = owner[lag]
    assert np.all(np.diff(own_sorted) >= 0)
    # per-patch interior index SETS equal the oracle's (LIndexSetData interior lists)
    for p in range(len(level.boxes)):
        mine = np.sort(lag[own_sorted == p])
        lst = ref["patches"][p]
        theirs = np.sort(lst["all_idx"][lst["interior_mask"]])
        assert np.array_equal(mine, theirs)
    # the index sets themselves, bit for bit (LIndexSetData::cacheLocalIndices, a11): every entry of the patch's ghost box
    # in the reference's order, periodic shifts, interior / ghost
    for p in range(len(level.boxes)):
        idx, sh, interior = ib.getPatchLists(p)
        lst = ref["patches"][p]
        assert np.array_equal(idx, lst["all_idx"])
        assert np.array_equal(sh.reshape(-1), lst["all_shift"])
        assert np.array_equal(interior, lst["interior_mask"])
    # same cell => ascending Lagrangian index (LDataManager.cpp:1505)
    key = [tuple(c) for c in cells[lag]]
    for i in range(1, N):
        if key[i] == key[i - 1] and own_sorted[i] == own_sorted[i - 1]:
            assert lag[i] > lag[i - 1]
    ib.close()


@pytest.mark.parametrize("ndim,split", [(2, (2, 2, 1)), (3, (2, 2, 1)), (3, (1, 1, 1))])
@pytest.mark.parametrize("kernel", ["IB_4", "IB_6", "BSPLINE_3", "IB_3", "BSPLINE_5"])
def test_resident_level_vs_reference_model(api, kernel, ndim, split):
    """spreadForce / interpolateVelocity on a periodic multi-patch level against the REFERENCE's
    model of the same operation: redundant spreading from each patch's ghost-box list, interiors
    kept (LDataManager.cpp:623-663); interpolation at the interior list after a ghost fill (:744-802)."""
    n = 32
    level = _level_case(ndim, n, split, kernel)
    g = level.gcw[0]
    N = 6000
    X = np.stack([_uniform(61 + d, N, -0.2, 1.3) for d in range(ndim)], axis=1)  # some outside: wrapped by rebin
    F = np.stack([_uniform(71 + d, N, -1.0, 1.0) for d in range(ndim)], axis=1)
    Xw, _ = orc.wrap_positions(X, level.x_lower, level.x_upper, level.periodic)
    Xw = Xw.reshape(-1, ndim)
    ref = orc.bin_level(level, Xw)
    ib = api.IBMethodB200(ndim, (0,) * ndim, (n - 1,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim, level.boxes, gcw=g,
                          kernel_fcn=kernel)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    # global periodic velocity field -> per patch interiors uploaded, ghosts garbage (filled by the halo)
    U_ref = np.zeros((N, ndim))
    f0 = []
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        u_full = _periodic_side_fields(pg, 300)
        lst = ref["patches"][p]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, ndim)[lst["interior_mask"]]
        orc.side_interp(kernel, pg, u_full, Xw, ii, sh.reshape(-1), U_ref)
        fo = [np.zeros(pg.side_shape(a)) for a in range(ndim)]
        orc.side_spread(kernel, pg, fo, Xw, F, lst["all_idx"], lst["all_shift"])
        f0.append(fo)
        for a in range(ndim):
            garbage = u_full[a].copy()
            interior = tuple(slice(g, s - g) for s in garbage.shape)
            mask = np.ones(garbage.shape, bool)
            mask[interior] = False
            garbage[mask] = 1e30  # must be overwritten by the ghost fill
            ib.grid_upload("u", p, a, garbage)
            ib.grid_upload("f", p, a, np.full(pg.side_shape(a), 0.25))  # f += ...: pre-existing content is kept
    ib.interpolateVelocity(fill_halo=True)
    U = ib.getLData("U")
    assert relerr(U, U_ref) <= TOL
    ib.spreadForce(accumulate_halo=True)
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        for a in range(ndim):
            got = ib.grid_download("f", p, a)
            interior = tuple(slice(g, s - g) for s in got.shape)
            assert relerr(got[interior] - 0.25, f0[p][a][interior]) <= TOL, (p, a)
    ib.close()


def _periodic_side_fields(pg, seed):
    """Smooth 1-periodic field + noise keyed by the GLOBAL periodic side index, so that every copy of a
    DOF (ghosts, shared faces, periodic images) carries the same value."""
    out = []
    ncell = [int(round(1.0 / pg.dx[d])) for d in range(pg.ndim)]
    for axis in range(pg.ndim):
        c = pg.side_coords(axis)
        f = np.sin(2 * np.pi * c[axis]) * np.cos(2 * np.pi * c[(axis + 1) % pg.ndim])
        gi = []
        for d in range(pg.ndim):
            cnt = pg.upper[d] - pg.lower[d] + 1 + (1 if d == axis else 0) + 2 * pg.gcw[d]
            gi.append(np.mod(np.arange(cnt) + pg.lower[d] - pg.gcw[d], ncell[d]))
        mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
        lin = np.zeros(f.shape, dtype=np.int64)
        mul = 1
        for d in range(pg.ndim):
            lin += mesh[d] * mul
            mul *= ncell[d]
        out.append(np.ascontiguousarray(f + 1e-3 * splitmix64_unit(seed + axis, lin.reshape(-1)).reshape(f.shape)))
    return out


def test_spread_is_bit_reproducible(api):
    """No floating-point atomics: two runs give identical bits, also after re-uploading the markers in
    a different storage order (the sums are ordered by (cell, Lagrangian index) only)."""
    ndim, n, kernel = 3, 48, "IB_4"
    level = _level_case(ndim, n, (1, 1, 1), kernel)
    N = 40000
    X = np.stack([_uniform(81 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    X[:20000] = 0.5 + 0.02 * (X[:20000] - 0.5)  # a dense cluster: many markers per cell
    F = np.stack([_uniform(91 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    runs = []
    for rep in range(3):
        ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1,) * 3, level.boxes, kernel_fcn=kernel)
        ib.setPositions(X)
        ib.setLData("F", F)
        ib.beginDataRedistribution()
        if rep == 2:
            ib.beginDataRedistribution()  # rebin from the already sorted storage order
        ib.spreadForce(accumulate_halo=True)
        runs.append([ib.grid_download("f", 0, a) for a in range(3)])
        ib.close()
    for a in range(3):
        assert np.array_equal(runs[0][a], runs[1][a])
        assert np.array_equal(runs[0][a], runs[2][a])


@pytest.mark.parametrize("kernel", ["IB_4", "IB_6", "PIECEWISE_LINEAR"])
def test_interior_plus_boundary_tiles_equal_the_whole(api, kernel):
    """ibk_spread_force_part / ibk_interpolate_velocity_part: parts 1 and 2 partition the marker tiles (both
    non-empty here), interior tiles touch neither ghost cells nor boundary faces."""
    ndim, n = 3, 96
    level = _level_case(ndim, n, (1, 1, 1), kernel)
    g = level.gcw[0]
    N = 120000
    X = np.stack([_uniform(301 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(311 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1,) * 3, level.boxes, kernel_fcn=kernel)
    pg = level.patch_geom(0)
    u = _periodic_side_fields(pg, 700)
    for a in range(3):
        ib.grid_upload("u", 0, a, u[a])
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    # reference: the whole operation without halo handling
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=False)
    f_all = [ib.grid_download("f", 0, a) for a in range(3)]
    ib.interpolateVelocity(fill_halo=False)
    U_all = ib.getLData("U")
    # interior tiles only: nothing lands in a ghost cell or on a boundary face, and something is spread
    ib.grid_fill("f", 0.0)
    ib.spreadForcePart(1)
    f_int = [ib.grid_download("f", 0, a) for a in range(3)]
    for a in range(3):
        inner = tuple(slice(g + 1, s - g - 1) for s in f_int[a].shape)
        outer = f_int[a].copy()
        outer[inner] = 0.0
        assert np.all(outer == 0.0) and np.any(f_int[a] != 0.0)
    ib.spreadForcePart(2)
    f_sum = [ib.grid_download("f", 0, a) for a in range(3)]
    for a in range(3):
        assert np.any(f_sum[a] != f_int[a])
        assert np.max(np.abs(f_sum[a] - f_all[a])) <= 1e-13 * np.max(np.abs(f_all[a]))
    ib.setLData("U", np.full((N, 3), 7.0))
    ib.interpolateVelocityPart(1)
    U1 = ib.getLData("U")
    ib.interpolateVelocityPart(2)
    U2 = ib.getLData("U")
    touched1 = np.any(U1 != 7.0, axis=1)
    assert 0 < touched1.sum() < N
    assert np.array_equal(U2, U_all)
    assert np.array_equal(U1[touched1], U_all[touched1])
    ib.close()


def test_marker_migration_between_two_contexts(api):
    """ibk_markers_set_ids / ibk_migrate_plan / _pack / _unpack with two contexts on one device standing for two
    ranks (the all-to-all is a pair of device copies): afterwards each side holds exactly the markers whose cell
    lies in its patch, rows in ascending global index, X/U/F intact, and the binning agrees with the oracle
    (LDataManager.cpp:1475-1476 ownership, :1824-1837 scatter)."""
    import torch
    from ibamr_b200 import halo
    ndim, n = 3, 16
    patches = halo.cartesian_patches(3, (2, 1, 1), (n, n, n))
    dom, xup = (2 * n, n, n), (2.0, 1.0, 1.0)
    level = orc.Level(3, (0,) * 3, dom, (0.0,) * 3, xup, (1, 1, 1), [(p.lower, p.upper) for p in patches], (3,) * 3)
    N = 30000
    ids = np.arange(N)
    X = np.stack([xup[d] * _uniform(141 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    U = np.stack([_uniform(151 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(161 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ref = orc.bin_level(level, X)
    ibs, start = [], []
    for r in range(2):
        ib = api.IBMethodB200(3, (0,) * 3, tuple(d - 1 for d in dom), (0.0,) * 3, xup, (1, 1, 1),
                              [(patches[r].lower, patches[r].upper)], kernel_fcn="IB_4", ctx=api.Context(0))  # one ctx per "rank"
        st = ids[r::2] if r == 0 else ids[1::2][::-1].copy()  # rank 1 starts in descending index order
        ib.setPositions(X[st])
        ib.setLData("U", U[st])
        ib.setLData("F", F[st])
        ib.setIds(st, N)
        ib.beginDataRedistribution()
        ibs.append(ib)
        start.append(st)
    lo = [p.lower for p in patches]
    hi = [p.upper for p in patches]
    # a marker that lies in no patch of the level is reported like an escaped point
    with pytest.raises(api.IBKError) as e:
        ibs[0].migrate_plan([lo[0]], [hi[0]], [0], 1, 0)
    assert e.value.code == api.IBK_ERR_ESCAPED
    counts = [ibs[r].migrate_plan(lo, hi, [0, 1], 2, r) for r in range(2)]
    assert counts[0][0] == 0 and counts[1][1] == 0 and counts[0][1] > 0 and counts[1][0] > 0
    bufs = [torch.zeros(max(int(counts[r].sum()), 1) * 10, dtype=torch.float64, device="cuda") for r in range(2)]
    torch.cuda.synchronize()
    for r in range(2):
        ibs[r].migrate_pack(bufs[r].data_ptr())
        ibs[r].ctx.synchronize()
    for r in range(2):
        ibs[r].migrate_unpack(bufs[1 - r].data_ptr(), int(counts[1 - r][r]), N)
        ibs[r].beginDataRedistribution()
    for r in range(2):
        mine = np.nonzero(ref["owner"] == r)[0]
        assert ibs[r].n_markers == len(mine)
        assert np.array_equal(ibs[r].getIds(), mine)
        assert np.array_equal(ibs[r].getLData("X"), X[mine])
        assert np.array_equal(ibs[r].getLData("U"), U[mine])
        assert np.array_equal(ibs[r].getLData("F"), F[mine])
        cells, owner = ibs[r].getCellsAndOwners()
        assert np.array_equal(cells, ref["cells"][mine]) and np.all(owner == 0)
        # sorted order: (cell in canonical order, then global index), as the oracle lists the patch
        order = ibs[r].getSortedLagrangianIndices()
        assert sorted(order.tolist()) == mine.tolist()
        # nothing more to move
        assert ibs[r].migrate_plan(lo, hi, [0, 1], 2, r).sum() == 0
        ibs[r].migrate_unpack(0, 0, N)
    for ib in ibs:
        ib.close()


def test_async_transfers_match_synchronous_ones(api):
    """ibk_grid_upload_async / ibk_grid_download_async (copy streams, device-side ordering) give the bits of
    the synchronous path: u uploaded while the spread runs, f downloaded while the interpolation runs."""
    import torch
    ndim, n, kernel = 3, 64, "IB_4"
    level = _level_case(ndim, n, (2, 1, 1), kernel)
    N = 60000
    X = np.stack([_uniform(181 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(191 + d, N, -1.0, 1.0) for d in range(3)], axis=1)

    def make():
        ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1,) * 3, level.boxes, kernel_fcn=kernel)
        ib.setPositions(X)
        ib.setLData("F", F)
        ib.beginDataRedistribution()
        return ib

    P = len(level.boxes)
    ref = make()
    u = [[np.ascontiguousarray(_uniform(7 + 10 * p + a, int(np.prod(ref.side_shape(p, a))), -1.0, 1.0)
                               .reshape(ref.side_shape(p, a))) for a in range(3)] for p in range(P)]
    for p in range(P):
        for a in range(3):
            ref.grid_upload("u", p, a, u[p][a])
    ref.spreadForce(accumulate_halo=True)
    ref.interpolateVelocity(fill_halo=True)
    f_ref = [[ref.grid_download("f", p, a) for a in range(3)] for p in range(P)]
    U_ref = ref.getLData("U")
    ref.close()

    ib = make()
    for rep in range(2):  # the second pass re-uses the arrays while transfers of the first could still be in flight
        hu = [[torch.from_numpy(u[p][a]).pin_memory() for a in range(3)] for p in range(P)]
        hf = [[torch.zeros(ib.side_shape(p, a), dtype=torch.float64).pin_memory() for a in range(3)] for p in range(P)]
        for p in range(P):
            for a in range(3):
                ib.grid_upload_async("u", p, a, hu[p][a].numpy())
        ib.grid_fill("f", 0.0)
        ib.spreadForce(accumulate_halo=True)
        for p in range(P):
            for a in range(3):
                ib.grid_download_async("f", p, a, hf[p][a].numpy())
        ib.interpolateVelocity(fill_halo=True)
        U = ib.getLData("U")
        ib.transfers_wait()
        assert np.array_equal(U, U_ref)
        for p in range(P):
            for a in range(3):
                assert np.array_equal(hf[p][a].numpy(), f_ref[p][a])
    ib.close()


@pytest.mark.parametrize("kernel", ["IB_4", "IB_6"])
def test_adjointness_and_moments_large(api, kernel):
    """Size-independent properties on a larger case (oracle not needed):
    <S F, u> h^3 = <F, J u> (spread and interpolate are adjoint), sum of spread force = sum of F
    (partition of unity), interpolation reproduces constants."""
    ndim, n = 3, 128
    level = _level_case(ndim, n, (1, 1, 1), kernel)
    g = level.gcw[0]
    N = 300000
    X = np.stack([_uniform(101 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(111 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1,) * 3, level.boxes, kernel_fcn=kernel)
    pg = level.patch_geom(0)
    u = _periodic_side_fields(pg, 400)
    for a in range(3):
        ib.grid_upload("u", 0, a, u[a])
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    ib.interpolateVelocity(fill_halo=True)
    U = ib.getLData("U")
    ib.spreadForce(accumulate_halo=True)
    h3 = (1.0 / n) ** 3
    lhs = 0.0
    for a in range(3):
        f = ib.grid_download("f", 0, a)
        sl = [slice(g, s - g) for s in f.shape]
        sl[2 - a] = slice(g, f.shape[2 - a] - g - 1)  # drop the periodic duplicate of side 0 along the axis
        lhs += np.sum(f[tuple(sl)] * u[a][tuple(sl)]) * h3
        assert abs(np.sum(f[tuple(sl)]) * h3 - np.sum(F[:, a])) <= 1e-9 * N ** 0.5
    rhs = np.sum(F * U)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(rhs), 1.0)
    for a in range(3):
        ib.grid_upload("u", 0, a, np.full(pg.side_shape(a), 3.25))
    ib.interpolateVelocity(fill_halo=False)
    assert np.max(np.abs(ib.getLData("U") - 3.25)) <= 1e-12
    ib.close()


@pytest.mark.parametrize("world,kernel,mode", [(2, "IB_4", "plain"), (2, "IB_4", "overlap"), (2, "IB_6", "overlap"), (4, "IB_4", "overlap"),
                                               (2, "IB_4", "migrate"), (2, "IB_4", "pipelined"), (4, "IB_6", "pipelined"), (8, "IB_4", "pipelined"),
                                               (8, "IB_4", "overlap")])
def test_ranks_as_contexts_of_one_process(api, world, kernel, mode):
    """The multi-rank path on ONE GPU (N > 1 had no driver-side parity evidence): `world` contexts of this
    process are the ranks of a loopback communicator of libibk.so (ibk_comm_init_loopback); each owns one patch of a
    2 x 1 x 1 / 2 x 2 x 1 decomposition of a periodic level and the markers in it.  The C-ABI plan, pack, message,
    unpack(-add) and migration code is the code the NCCL transport runs; the overlapped sequence is the one bench.py times
    at N > 1.  Compared per rank with the oracle's model of the reference (redundant ghost-region spreading)."""
    from ibamr_b200 import halo
    n = 64 if mode == "overlap" else 32  # with 64 cells per rank there are interior tiles
    pgrid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]  # 8: the process grid of the 8-GPU benchmark (every face remote)
    patches = halo.cartesian_patches(3, pgrid, (n, n, n))
    dom = tuple(n * pgrid[d] for d in range(3))
    g = orc.min_ghost_width(kernel)
    xup = tuple(float(p) for p in pgrid)
    level = orc.Level(3, (0,) * 3, dom, (0.0,) * 3, xup, (1, 1, 1), [(p.lower, p.upper) for p in patches], (g,) * 3)
    N = 40000
    X = np.stack([xup[d] * _uniform(61 + d, N, 0.0, 1.0) for d in range(3)], axis=1)
    F = np.stack([_uniform(71 + d, N, -1.0, 1.0) for d in range(3)], axis=1)
    ref = orc.bin_level(level, X)
    ctxs = [api.Context(0) for _ in range(world)]
    halo.CommExchange.init_loopback(ctxs)
    ibs, hxs, us = [], [], []
    for r in range(world):
        me = patches[r]
        ib = api.IBMethodB200(3, (0,) * 3, tuple(d - 1 for d in dom), (0.0,) * 3, xup, (1, 1, 1), [(me.lower, me.upper)], gcw=g,
                              kernel_fcn=kernel, ctx=ctxs[r])
        pg = level.patch_geom(r)
        u = _periodic_side_fields_global(pg, dom, 300)
        for a in range(3):
            garbage = u[a].copy()
            mask = np.ones(garbage.shape, bool)
            mask[tuple(slice(g, s - g) for s in garbage.shape)] = False
            garbage[mask] = 1e30  # the ghost values must come from the exchange
            ib.grid_upload("u", 0, a, garbage)
            ib.grid_upload("f", 0, a, np.full(pg.side_shape(a), 0.25))
        ibs.append(ib)
        us.append(u)
        hxs.append(halo.CommExchange(ib, patches))
    mine = [np.nonzero(ref["owner"] == r)[0] for r in range(world)]
    if mode == "migrate":
        import ctypes as C
        for r in range(world):  # start from an arbitrary distribution (index mod world) and let the markers find their owners
            start = np.arange(r, N, world)
            ibs[r].setPositions(X[start])
            ibs[r].setLData("F", F[start])
            ibs[r].setIds(start, N)
            ibs[r].beginDataRedistribution()
        arr = (C.c_void_p * world)(*[c.h for c in ctxs])
        moved = C.c_int()
        ctxs[0].check(ctxs[0].lib.ibk_migrate_loopback(arr, world, C.c_uint(N), C.byref(moved)))
        assert moved.value > 0
        for r in range(world):
            ibs[r].n_markers = int(ctxs[r].lib.ibk_markers_count(ctxs[r].h))
            ibs[r].beginDataRedistribution()
            assert np.array_equal(ibs[r].getIds(), mine[r]), "after the migration a rank holds exactly the markers of its patch"
            assert np.array_equal(ibs[r].getLData("X"), X[mine[r]]) and np.array_equal(ibs[r].getLData("F"), F[mine[r]])
    else:
        for r in range(world):
            ibs[r].setPositions(X[mine[r]])
            ibs[r].setLData("F", F[mine[r]])
            ibs[r].beginDataRedistribution()
    chk = lambda r, rc: ctxs[r].check(rc)
    if mode == "overlap":
        for r in range(world):
            chk(r, ctxs[r].lib.ibk_spread_begin(ctxs[r].h))
            ibs[r].spreadForcePart(2)
            hxs[r].accumulate_post()
            ibs[r].spreadForcePart(1)
            ibs[r].halo("f")
        for r in range(world):
            hxs[r].accumulate_finish()
            chk(r, ctxs[r].lib.ibk_spread_end(ctxs[r].h))
        for r in range(world):
            hxs[r].fill_post()
            ibs[r].halo("u")
            ibs[r].interpolateVelocityPart(1)
        for r in range(world):
            hxs[r].fill_finish()
            ibs[r].interpolateVelocityPart(2)
    elif mode == "pipelined":
        # the sequence bench.py times at N > 1: the u ghosts travel during the spread, the f ghost contributions during the
        # interpolation; neither kernel is split
        for r in range(world):
            hxs[r].fill_post()
        for r in range(world):
            chk(r, ctxs[r].lib.ibk_spread_begin(ctxs[r].h))
            ibs[r].spreadForce(accumulate_halo=False)
            hxs[r].accumulate_post()
            ibs[r].halo("f")
        for r in range(world):
            ibs[r].halo("u")
            hxs[r].fill_finish()
            ibs[r].interpolateVelocity(fill_halo=False)
        for r in range(world):
            hxs[r].accumulate_finish()
            chk(r, ctxs[r].lib.ibk_spread_end(ctxs[r].h))
    else:
        for r in range(world):
            chk(r, ctxs[r].lib.ibk_spread_begin(ctxs[r].h))
            ibs[r].spreadForce(accumulate_halo=False)
            hxs[r].accumulate_post()
            ibs[r].halo("f")
        for r in range(world):
            hxs[r].accumulate_finish()
            chk(r, ctxs[r].lib.ibk_spread_end(ctxs[r].h))
        for r in range(world):
            ibs[r].halo("u")
            hxs[r].fill_post()
        for r in range(world):
            hxs[r].fill_finish()
            ibs[r].interpolateVelocity(fill_halo=False)
    for r in range(world):
        pg = level.patch_geom(r)
        U = ibs[r].getLData("U")
        lst = ref["patches"][r]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, 3)[lst["interior_mask"]]
        U_ref = orc.side_interp(kernel, pg, us[r], X, ii, sh.reshape(-1))
        f_ref = [np.zeros(pg.side_shape(a)) for a in range(3)]
        orc.side_spread(kernel, pg, f_ref, X, F, lst["all_idx"], lst["all_shift"])
        assert relerr(U, U_ref[mine[r]]) <= TOL
        for a in range(3):
            f = ibs[r].grid_download("f", 0, a)
            sl = tuple(slice(g, s - g) for s in f.shape)
            assert np.max(np.abs(f[sl] - 0.25 - f_ref[a][sl])) <= TOL * np.max(np.abs(f_ref[a][sl]))
        assert hxs[r].bytes(0) > 0 and hxs[r].bytes(1) > 0
    for ib in ibs:
        ib.close()


def _periodic_side_fields_global(pg, ncell, seed):
    """Smooth + noisy side-centred fields that are periodic over the GLOBAL domain (the same value at periodic images)."""
    out = []
    for axis in range(3):
        c = pg.side_coords(axis)
        f = np.sin(2 * np.pi * c[axis] / (ncell[axis] * pg.dx[axis])) * np.cos(2 * np.pi * c[(axis + 1) % 3] / (ncell[(axis + 1) % 3] * pg.dx[0]))
        gi = []
        for d in range(3):
            cnt = pg.upper[d] - pg.lower[d] + 1 + (1 if d == axis else 0) + 2 * pg.gcw[d]
            gi.append(np.mod(np.arange(cnt) + pg.lower[d] - pg.gcw[d], ncell[d]))
        mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
        lin = mesh[0] + ncell[0] * (mesh[1] + ncell[1] * mesh[2])
        out.append(np.ascontiguousarray(f + 1e-3 * splitmix64_unit(seed + axis, lin.reshape(-1)).reshape(f.shape)))
    return out


def test_rebin_escape_check_leaves_positions_untouched_and_owned_count(api):
    """IBMethod's error_if_points_leave_domain (IBMethod.cpp:2060 -> LDataManager.cpp:1410-1416): the reference aborts BEFORE it
    moves anything, so a refused re-bin leaves X as it was; without the flag the points are clamped.  ibk_markers_owned_count
    tells how many markers a local patch accepted (here one patch covers the lower half of the domain only)."""
    import ctypes as C
    n = 16
    ib = api.IBMethodB200(3, (0,) * 3, (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (0, 0, 0), [((0, 0, 0), (n // 2 - 1, n - 1, n - 1))],
                          kernel_fcn="IB_4", ctx=api.Context(0), error_if_points_leave_domain=True)
    N = 1000
    X = np.stack([_uniform(171 + d, N, 0.01, 0.99) for d in range(3)], axis=1)
    X[7, 1] = 1.25  # outside the (non-periodic) domain
    ib.setPositions(X)
    with pytest.raises(api.IBKError) as e:
        ib.beginDataRedistribution()
    assert e.value.code == api.IBK_ERR_ESCAPED
    assert np.array_equal(ib.getLData("X"), X)
    ib.error_if_points_leave_domain = False
    ib.beginDataRedistribution()
    Xc = ib.getLData("X")
    assert Xc[7, 1] < 1.0 and np.array_equal(np.delete(Xc, 7, axis=0), np.delete(X, 7, axis=0))
    owned = C.c_int(-1)
    ib.ctx.check(ib.ctx.lib.ibk_markers_owned_count(ib.ctx.h, C.byref(owned)))
    assert owned.value == int(np.sum(np.floor(Xc[:, 0] * n) < n // 2))
    ib.close()


@pytest.mark.parametrize("kernel", ["IB_4", "PIECEWISE_LINEAR"])
def test_position_only_interpolate_far_box(api, kernel):
    """a position-only interpolate whose box reaches farther from the patch than the binning accepts (gcw + 4 cells).
    The reference lists every marker of the box (LEInteractor::buildLocalIndices, LEInteractor.cpp:6088-6126); one whose stencil
    finds no array point gets 0 (:3117-3120); markers outside the box keep the caller's values."""
    ndim, n = 2, 8
    g = orc.min_ghost_width(kernel)
    lo, hi = (0, 0), (n - 1, n - 1)
    dx = (1.0 / n,) * 2
    pg = orc.PatchGeom(lo, hi, (0.0, 0.0), (1.0, 1.0), dx, (g,) * 2)
    box = api.Box(lo, hi)
    patch = api.Patch(box, (0.0, 0.0), (1.0, 1.0), dx)
    u = _side_fields(pg, 100)
    N = 4000
    X = np.stack([_uniform(3 + d, N, -2.5, 3.5) for d in range(2)], axis=1)  # up to 20 cells away from the patch
    far = api.Box((-14, -14), (n - 1 + 14, n - 1 + 14))
    Q0 = np.full((N, 2), 5.0)
    idx = orc.indices_in_box(X, pg, far.lower, far.upper)
    Qo = orc.side_interp(kernel, pg, u, X, idx, None, Q0.copy())
    q = api.SideData(box, 1, g, u)
    Q = Q0.copy()
    api.LEInteractor.interpolate(Q, 2, X, 2, q, patch, far, kernel)
    listed = np.zeros(N, bool)
    listed[idx] = True
    assert listed.sum() > 100 and (~listed).sum() > 100
    assert np.array_equal(Q[~listed], Q0[~listed])
    assert np.count_nonzero(Qo[listed].any(axis=1) == 0) > 50  # the far ones interpolate to exactly 0
    assert np.max(np.abs(Q - Qo)) <= TOL * np.max(np.abs(Qo))
