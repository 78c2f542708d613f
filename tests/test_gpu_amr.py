"""N3 of SURVEY 8(f), AMR part: conservative-linear prolongation of f and conservative coarsening of u between two levels
resident on one device (one context per level), against the oracle's restatement of SAMRAI's two operators (bit for bit:
both sides evaluate the same expressions without contraction), and a two-level spread / interpolation in which the coarse
level holds markers too (LDataManager::spread :606-667, ::interp :728-813)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import splitmix64_unit

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def api():
    from ibamr_b200 import api as _api
    _api.default_context()
    return _api


def _levels(ndim, nc, ratio, g, fine_boxes, coarse_boxes=None):
    coarse = orc.Level(ndim, (0,) * ndim, (nc,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim,
                       coarse_boxes or [((0,) * ndim, (nc - 1,) * ndim)], (g,) * ndim)
    fine = orc.Level(ndim, (0,) * ndim, (nc * ratio,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim, fine_boxes, (g,) * ndim)
    return coarse, fine


def _field(level, seed, smooth=True):
    """analytic + noise on every patch array, ghosts included, consistent between patches and across the periodic wrap"""
    ndim = level.ndim
    out = []
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        arrs = []
        for a in range(ndim):
            c = pg.side_coords(a)
            f = np.sin(2 * np.pi * c[a]) * np.cos(2 * np.pi * c[(a + 1) % ndim]) if smooth else np.zeros(pg.side_shape(a))
            gi = [np.mod(np.arange(pg.upper[d] - pg.lower[d] + 1 + (1 if d == a else 0) + 2 * pg.gcw[d]) + pg.lower[d] - pg.gcw[d],
                         level.domain_ncells[d]) for d in range(ndim)]
            mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
            lin = sum(mesh[d] * int(np.prod(level.domain_ncells[:d])) for d in range(ndim))
            arrs.append(np.ascontiguousarray(np.broadcast_to(f, pg.side_shape(a)) + 0.3 * splitmix64_unit(seed + a, lin.reshape(-1)).reshape(pg.side_shape(a))))
        out.append(arrs)
    return out


def _make(api, level, kernel, g):
    return api.IBMethodB200(level.ndim, level.domain_lower, level.domain_upper(), level.x_lower, level.x_upper, level.periodic,
                            level.boxes, gcw=g, kernel_fcn=kernel, ctx=api.Context(0))


def _fine_boxes(ndim, lo, n, split):
    """a cube [lo, lo + n) cut into split^ndim patches"""
    w = n // split
    boxes = []
    for k in np.ndindex(*(split,) * ndim):
        l = tuple(lo + w * k[ndim - 1 - d] for d in range(ndim))
        boxes.append((l, tuple(x + w - 1 for x in l)))
    return boxes


@pytest.mark.parametrize("ndim,nc,ratio,g", [(2, 16, 4, 3), (3, 12, 4, 3), (3, 24, 2, 3), (2, 12, 3, 2)])
def test_refine_and_coarsen_bit_exact(api, ndim, nc, ratio, g):
    lo = (nc // 4) * ratio
    n = (nc // 2) * ratio
    coarse_boxes = None
    if ndim == 3 and ratio == 2:  # two coarse patches: the refine reads across their common face through the ghosts
        coarse_boxes = [((0, 0, 0), (nc // 2 - 1, nc - 1, nc - 1)), ((nc // 2, 0, 0), (nc - 1, nc - 1, nc - 1))]
    coarse, fine = _levels(ndim, nc, ratio, g, _fine_boxes(ndim, lo, n, 1 if n < 24 else 2), coarse_boxes)
    r = (ratio,) * ndim
    ibc, ibf = _make(api, coarse, "IB_4", g), _make(api, fine, "IB_4", g)
    try:
        # ---- prolongation of f
        Cf = _field(coarse, 100)
        Ff = _field(fine, 200, smooth=False)
        for p in range(len(coarse.boxes)):
            for a in range(ndim):
                ibc.grid_upload("f", p, a, Cf[p][a])
        for p in range(len(fine.boxes)):
            for a in range(ndim):
                ibf.grid_upload("f", p, a, Ff[p][a])
        n_ref = orc.amr_refine_side(coarse, fine, r, Cf, Ff)
        assert ibf.prolongFrom(ibc, r, "f") == n_ref
        for p in range(len(fine.boxes)):
            for a in range(ndim):
                assert np.array_equal(ibf.grid_download("f", p, a), Ff[p][a]), (p, a)
        # ---- coarsening of u
        Cu = _field(coarse, 300)
        Fu = _field(fine, 400)
        for p in range(len(coarse.boxes)):
            for a in range(ndim):
                ibc.grid_upload("u", p, a, Cu[p][a])
        for p in range(len(fine.boxes)):
            for a in range(ndim):
                ibf.grid_upload("u", p, a, Fu[p][a])
        n_ref = orc.amr_coarsen_side(coarse, fine, r, Cu, Fu)
        assert ibc.coarsenFrom(ibf, r, "u") == n_ref and n_ref > 0
        for p in range(len(coarse.boxes)):
            for a in range(ndim):
                assert np.array_equal(ibc.grid_download("u", p, a), Cu[p][a]), (p, a)
    finally:
        ibc.close()
        ibf.close()


def test_two_level_spread_and_interp_with_markers_on_both_levels(api):
    """Coarse 16^3 (one periodic patch) + ratio-4 fine level over the central 32^3 fine cells... the hierarchy of BASELINE's C4
    with markers on BOTH levels.  spread: coarse spread, ghost fill of the coarse f, prolongation, fine spread on top (f +=);
    interp: fine u coarsened onto the coarse level, ghost fill, coarse interpolation."""
    ndim, nc, ratio, g, kernel = 3, 16, 4, 3, "IB_4"
    coarse, fine = _levels(ndim, nc, ratio, g, _fine_boxes(3, 16, 32, 2))
    r = (ratio,) * 3
    Nc, Nf = 3000, 6000
    Xc = np.stack([splitmix64_unit(60 + d, np.arange(Nc)) for d in range(3)], axis=1)
    Fc = np.stack([2 * splitmix64_unit(1 + d, np.arange(Nc)) - 1 for d in range(3)], axis=1)
    Xf = 0.5 + 0.16 * (2 * np.stack([splitmix64_unit(70 + d, np.arange(Nf)) for d in range(3)], axis=1) - 1)  # >= gcw fine cells inside
    Ff = np.stack([2 * splitmix64_unit(11 + d, np.arange(Nf)) - 1 for d in range(3)], axis=1)
    ibc, ibf = _make(api, coarse, kernel, g), _make(api, fine, kernel, g)
    try:
        # ---------------- spread
        ibc.setPositions(Xc)
        ibc.setLData("F", Fc)
        ibc.beginDataRedistribution()
        ibf.setPositions(Xf)
        ibf.setLData("F", Ff)
        ibf.beginDataRedistribution()
        ibc.grid_fill("f", 0.0)
        ibc.spreadForce(accumulate_halo=True)
        # reference model on the coarse level: one periodic patch, all markers with their periodic images, interiors kept
        refc = orc.bin_level(coarse, Xc)
        pgc = coarse.patch_geom(0)
        fc = [np.zeros(pgc.side_shape(a)) for a in range(3)]
        orc.side_spread(kernel, pgc, fc, Xc, Fc, refc["patches"][0]["all_idx"], refc["patches"][0]["all_shift"])
        # the prolongation reads coarse ghosts: periodic copies of the interiors (what the refine schedule's coarse scratch holds)
        got_c = [ibc.grid_download("f", 0, a) for a in range(3)]
        for a in range(3):
            inner = tuple(slice(g, s - g) for s in fc[a].shape)
            assert np.max(np.abs(got_c[a][inner] - fc[a][inner])) / np.max(np.abs(fc[a])) <= TOL
        fcw = []
        for a in range(3):
            arr = got_c[a].copy()  # take the device's interiors so that the refine is compared bit for bit
            idx = [np.mod(np.arange(arr.shape[2 - d]) - g, nc) + g for d in range(3)]
            # the upper face of component a is the periodic image of the lower one: wrap with period nc on the side index too
            arr = arr[np.ix_(idx[2], idx[1], idx[0])]
            fcw.append(np.ascontiguousarray(arr))
            ibc.grid_upload("f", 0, a, fcw[a])
        ff = [[np.zeros(fine.patch_geom(p).side_shape(a)) for a in range(3)] for p in range(len(fine.boxes))]
        orc.amr_refine_side(coarse, fine, r, [fcw], ff)
        ibf.prolongFrom(ibc, r, "f")
        ibf.spreadForce(accumulate_halo=True)  # f += S[F] on top of the prolonged field
        reff = orc.bin_level(fine, Xf)
        for p in range(len(fine.boxes)):
            pg = fine.patch_geom(p)
            add = [np.zeros(pg.side_shape(a)) for a in range(3)]
            orc.side_spread(kernel, pg, add, Xf, Ff, reff["patches"][p]["all_idx"], reff["patches"][p]["all_shift"])
            for a in range(3):
                got = ibf.grid_download("f", p, a)
                inner = tuple(slice(g, s - g) for s in got.shape)
                want = ff[p][a] + add[a]
                assert np.max(np.abs(got[inner] - want[inner])) / np.max(np.abs(want[inner])) <= TOL, (p, a)
        # ---------------- interpolation: fine u -> coarse u under the fine region, then the coarse markers
        Cu, Fu = _field(coarse, 300), _field(fine, 400)
        for a in range(3):
            ibc.grid_upload("u", 0, a, Cu[0][a])
        for p in range(len(fine.boxes)):
            for a in range(3):
                ibf.grid_upload("u", p, a, Fu[p][a])
        orc.amr_coarsen_side(coarse, fine, r, Cu, Fu)
        ibc.coarsenFrom(ibf, r, "u")
        ibc.interpolateVelocity(fill_halo=True)
        U = ibc.getLData("U")
        # reference: ghost fill = periodic copy of the (coarsened) interiors, then the interpolation at the interior list
        cu = []
        for a in range(3):
            arr = Cu[0][a]
            idx = [np.mod(np.arange(arr.shape[2 - d]) - g, nc) + g for d in range(3)]
            cu.append(np.ascontiguousarray(arr[np.ix_(idx[2], idx[1], idx[0])]))
        U_ref = np.zeros((Nc, 3))
        lst = refc["patches"][0]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, 3)[lst["interior_mask"]]
        orc.side_interp(kernel, pgc, cu, Xc, ii, sh.reshape(-1), U_ref)
        assert np.max(np.abs(U - U_ref)) / np.max(np.abs(U_ref)) <= TOL
    finally:
        ibc.close()
        ibf.close()
