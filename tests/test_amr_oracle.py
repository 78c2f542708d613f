"""CPU checks of the oracle's restatement of the two AMR transfer operators around the path (N3 of SURVEY 8(f)):
SAMRAI's CONSERVATIVE_LINEAR_REFINE and CONSERVATIVE_COARSEN for side data (third party, not in the reference tree, so
there is no fixture to pin them on: parity UNPINNED).  What can be checked without the source are the properties the
operators are named after: linear fields are reproduced, coarsen(refine(c)) = c (conservation), no new extrema."""
import numpy as np
import pytest

from oracle import oracle as orc


def two_levels(ndim, nc=8, ratio=4, g=2):
    r = (ratio,) * ndim
    coarse = orc.Level(ndim, (0,) * ndim, (nc,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim,
                       [((0,) * ndim, (nc - 1,) * ndim)], (g,) * ndim)
    lo, hi = nc // 4 * ratio, 3 * nc // 4 * ratio - 1
    mid = (lo + hi + 1) // 2
    boxes = [((lo,) * ndim, (mid - 1,) + (hi,) * (ndim - 1)), ((mid,) + (lo,) * (ndim - 1), (hi,) * ndim)]
    fine = orc.Level(ndim, (0,) * ndim, (nc * ratio,) * ndim, (0.0,) * ndim, (1.0,) * ndim, (1,) * ndim, boxes, (g,) * ndim)
    return coarse, fine, r


def side_arrays(level, fn):
    out = []
    for p in range(len(level.boxes)):
        pg = level.patch_geom(p)
        out.append([np.ascontiguousarray(np.broadcast_to(fn(a, pg.side_coords(a)), pg.side_shape(a)), dtype=np.float64) for a in range(level.ndim)])
    return out


@pytest.mark.parametrize("ndim", [2, 3])
def test_refine_reproduces_linear_fields(ndim):
    coarse, fine, r = two_levels(ndim)
    coef = [0.3, -1.1, 0.7][:ndim]
    lin = lambda a, c: 2.0 + sum(coef[d] * c[d] for d in range(ndim)) + 0.25 * a
    C = side_arrays(coarse, lin)
    F = side_arrays(fine, lambda a, c: np.nan)
    n = orc.amr_refine_side(coarse, fine, r, C, F)
    assert n > 0
    ref = side_arrays(fine, lin)
    for p in range(len(fine.boxes)):
        for a in range(ndim):
            m = ~np.isnan(F[p][a])
            g = fine.gcw[0]
            inner = tuple(slice(g, s - g) for s in F[p][a].shape)
            assert m[inner].all()  # the whole patch interior is reached
            assert np.max(np.abs(F[p][a][m] - ref[p][a][m])) < 1e-13


@pytest.mark.parametrize("ndim", [2, 3])
def test_coarsen_of_refine_is_the_identity_and_no_new_extrema(ndim):
    coarse, fine, r = two_levels(ndim)
    rng = np.random.default_rng(5)
    C = [[rng.standard_normal(coarse.patch_geom(0).side_shape(a)) for a in range(ndim)]]
    F = side_arrays(fine, lambda a, c: 0.0)
    orc.amr_refine_side(coarse, fine, r, C, F)
    C2 = [[np.full_like(C[0][a], np.nan) for a in range(ndim)]]
    n = orc.amr_coarsen_side(coarse, fine, r, C2, F)
    assert n > 0
    for a in range(ndim):
        m = ~np.isnan(C2[0][a])
        assert m.any()
        assert np.max(np.abs(C2[0][a][m] - C[0][a][m])) < 1e-13
        # monotonised slopes: a refined value stays within the range of the coarse stencil it came from
        assert F[0][a].max() <= C[0][a].max() + 1e-12 and F[0][a].min() >= C[0][a].min() - 1e-12


def test_coarsen_is_the_area_weighted_mean():
    coarse, fine, r = two_levels(3)
    F = side_arrays(fine, lambda a, c: np.sin(5 * c[0]) * np.cos(3 * c[1]) + c[2] ** 2)
    C = [[np.full(coarse.patch_geom(0).side_shape(a), np.nan) for a in range(3)]]
    orc.amr_coarsen_side(coarse, fine, r, C, F)
    # coarse x-side (4, 3, 2) of component 0 <- the 16 fine x-sides at i = 16, j = 12..15, k = 8..11 of the patch that owns them
    g = coarse.gcw[0]
    got = C[0][0][2 + g, 3 + g, 4 + g]
    p = 0 if 16 <= fine.boxes[0][1][0] + 1 and fine.boxes[0][0][0] <= 16 else 1
    lo = fine.boxes[p][0]
    gf = fine.gcw[0]
    blk = F[p][0][8 - lo[2] + gf:12 - lo[2] + gf, 12 - lo[1] + gf:16 - lo[1] + gf, 16 - lo[0] + gf]
    assert abs(got - blk.mean()) < 1e-14
