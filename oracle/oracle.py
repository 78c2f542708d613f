"""ctypes front end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The arithmetic lives in oracle/le_kernels.c and oracle/le_driver.c
(each function there cites the reference file:line it restates); this file adds the small
numpy orchestration that the reference does with SAMRAI/PETSc objects:

  * PatchGeom / level helpers                     (SAMRAI CartesianPatchGeometry role)
  * side_* / cell_* position-only LEInteractor calls
        ibtk/src/lagrangian/LEInteractor.cpp:2805-2863 (CellData), :3045-3127, :4188-4263 (SideData)
  * ghost_accumulate                              SAMRAIGhostDataAccumulator::accumulateGhostData
        ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:295-353 (+ DOF sharing of faces,
        ibtk/src/math/PETScVecUtilities.cpp:519-611)
  * bin_level                                     LDataManager::beginDataRedistribution binning
        ibtk/src/lagrangian/LDataManager.cpp:1397-1508, LIndexSetData.cpp:53-141

Array convention: a Fortran array u(lo0-g:hi0+g, lo1-g:hi1+g[, lo2-g:hi2+g]) is held as a
C-ordered numpy array of shape ([n2,] n1, n0), i.e. the same memory.

Parity status: PINNED (tests/test_oracle_golden.py, fixtures under tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KERNELS = {"PIECEWISE_LINEAR": 0, "IB_4": 1, "IB_6": 2, "BSPLINE_3": 3, "BSPLINE_4": 4, "IB_3": 5, "BSPLINE_5": 6, "BSPLINE_6": 7,
           "PIECEWISE_CUBIC": 8, "IB_5": 9, "PIECEWISE_CONSTANT": 10,
           "COMPOSITE_BSPLINE_32": 11, "COMPOSITE_BSPLINE_23": 12, "COMPOSITE_BSPLINE_43": 13, "COMPOSITE_BSPLINE_34": 14,
           "COMPOSITE_BSPLINE_54": 15, "COMPOSITE_BSPLINE_45": 16, "COMPOSITE_BSPLINE_65": 17, "COMPOSITE_BSPLINE_56": 18,
           "DISCONTINUOUS_LINEAR": 19, "IB_4_W8": 20}
# LEInteractor::getStencilSize (LEInteractor.cpp:2052-2108) and getMinimumGhostWidth (:2110-2114)
STENCIL_SIZE = {"PIECEWISE_LINEAR": 2, "IB_4": 4, "IB_6": 6, "BSPLINE_3": 4, "BSPLINE_4": 4, "IB_3": 4, "BSPLINE_5": 6, "BSPLINE_6": 6,
                "PIECEWISE_CUBIC": 4, "IB_5": 6, "PIECEWISE_CONSTANT": 1,
                "COMPOSITE_BSPLINE_32": 4, "COMPOSITE_BSPLINE_23": 4, "COMPOSITE_BSPLINE_43": 4, "COMPOSITE_BSPLINE_34": 4,
                "COMPOSITE_BSPLINE_54": 5, "COMPOSITE_BSPLINE_45": 5, "COMPOSITE_BSPLINE_65": 6, "COMPOSITE_BSPLINE_56": 6,
                "DISCONTINUOUS_LINEAR": 2, "IB_4_W8": 8}  # LEInteractor::getStencilSize (LEInteractor.cpp:2052-2103)


def min_ghost_width(kernel: str) -> int:
    return int(np.floor(0.5 * STENCIL_SIZE[kernel])) + 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("le_kernels.c", "le_driver.c", "le_force.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.le_oracle_baseline_create.restype = C.c_void_p
        _LIB.le_oracle_baseline_f.restype = C.POINTER(C.c_double)
        _LIB.le_oracle_baseline_u.restype = C.POINTER(C.c_double)
        _LIB.le_oracle_baseline_array_size.restype = C.c_size_t
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------------------------
# raw funnel (seam B4): same argument meaning as the Fortran routines
# ----------------------------------------------------------------------------------------------
def interp_raw(kernel, ndim, dx, x_lower, depth, ilower, iupper, nugc, u, indices, Xshift, X, V=None, axis=0):
    """`axis`: the extra argument of the axis-dependent Fortran routines (composite B-splines, discontinuous linear)."""
    X = _f64(X)
    indices = _i32(indices)
    Xshift = _f64(Xshift)
    u = _f64(u)
    if V is None:
        V = np.zeros((X.size // ndim, depth))
    lib().le_oracle_interp(
        KERNELS[kernel] | (axis << 8), ndim, _dp(_f64(dx)), _dp(_f64(x_lower)), depth, _ip(_i32(ilower)), _ip(_i32(iupper)),
        _ip(_i32(nugc)), _dp(u), _ip(indices), _dp(Xshift), int(indices.size), _dp(X), _dp(V))
    return V


def spread_raw(kernel, ndim, dx, x_lower, depth, indices, Xshift, X, V, ilower, iupper, nugc, u, axis=0):
    assert u.dtype == np.float64 and u.flags.c_contiguous
    X = _f64(X)
    V = _f64(V)
    indices = _i32(indices)
    Xshift = _f64(Xshift)
    lib().le_oracle_spread(
        KERNELS[kernel] | (axis << 8), ndim, _dp(_f64(dx)), _dp(_f64(x_lower)), depth, _ip(indices), _dp(Xshift),
        int(indices.size), _dp(X), _dp(V), _ip(_i32(ilower)), _ip(_i32(iupper)), _ip(_i32(nugc)), _dp(u))
    return u


def ib4_kernel_fcn(r):
    """the default of LEInteractor::s_kernel_fcn (the 4-point function, LEInteractor.cpp:1526-1546), scalar"""
    a = abs(r)
    if a >= 2.0:
        return 0.0
    a2 = a * a
    if a < 1.0:
        return -a / 4.0 + 3.0 / 8.0 + np.sqrt(-4.0 * a2 + 4.0 * a + 1.0) / 8.0
    return -a / 4.0 + 5.0 / 8.0 - np.sqrt(12.0 * a - 7.0 - 4.0 * a2) / 8.0


def _user_stencils(kernel_fcn, stencil, ndim, dx, x_lower, ilower, iupper, nugc, X, Xshift, l, s):
    """stencil range and weights of list entry l (marker s), LEInteractor.cpp:6158-6226 / 6286-6354"""
    lo, hi, w = [], [], []
    for d in range(ndim):
        xs = X[s * ndim + d] + (Xshift[l * ndim + d] if Xshift is not None else 0.0)
        center = int(np.floor((xs - x_lower[d]) / dx[d])) + ilower[d]
        x_cell = x_lower[d] + (float(center - ilower[d]) + 0.5) * dx[d]
        if stencil % 2 == 0:
            if X[s * ndim + d] < x_cell:  # (sic: the unshifted position)
                a, b = center - stencil // 2, center + stencil // 2 - 1
            else:
                a, b = center - stencil // 2 + 1, center + stencil // 2
        else:
            a, b = center - stencil // 2, center + stencil // 2
        a = min(max(a, ilower[d] - nugc[d]), iupper[d] + nugc[d])
        b = min(max(b, ilower[d] - nugc[d]), iupper[d] + nugc[d])
        lo.append(a)
        hi.append(b)
        w.append([kernel_fcn((xs - (x_cell + float(ic - center) * dx[d])) / dx[d]) for ic in range(a, b + 1)])
    return lo, hi, w


def user_interp_raw(kernel_fcn, stencil, ndim, dx, x_lower, depth, ilower, iupper, nugc, u, indices, Xshift, X, V):
    """LEInteractor::userDefinedInterpolate (LEInteractor.cpp:6128-6257).  u: [depth][z][y][x] with ghosts, V: [n][depth]."""
    X, V = np.asarray(X, dtype=np.float64).reshape(-1), V.reshape(-1, depth)
    sh = None if Xshift is None else np.asarray(Xshift, dtype=np.float64).reshape(-1)
    shape = tuple(iupper[d] - ilower[d] + 1 + 2 * nugc[d] for d in range(ndim))[::-1]
    q = np.asarray(u, dtype=np.float64).reshape((depth,) + shape)
    for l, s in enumerate(np.asarray(indices).reshape(-1)):
        lo, hi, w = _user_stencils(kernel_fcn, stencil, ndim, dx, x_lower, ilower, iupper, nugc, X, sh, l, int(s))
        for d in range(depth):
            acc = 0.0
            for i2 in range(lo[2], hi[2] + 1) if ndim == 3 else [0]:
                for i1 in range(lo[1], hi[1] + 1):
                    for i0 in range(lo[0], hi[0] + 1):
                        j = (i0 - ilower[0] + nugc[0], i1 - ilower[1] + nugc[1]) + ((i2 - ilower[2] + nugc[2],) if ndim == 3 else ())
                        ww = w[0][i0 - lo[0]] * w[1][i1 - lo[1]]
                        if ndim == 3:
                            ww = ww * w[2][i2 - lo[2]]
                        acc += ww * q[(d,) + j[::-1]]
            V[int(s), d] = acc
    return V


def user_spread_raw(kernel_fcn, stencil, ndim, dx, x_lower, depth, indices, Xshift, X, V, ilower, iupper, nugc, u):
    """LEInteractor::userDefinedSpread (LEInteractor.cpp:6259-6382): serial over the list, q += w0 w1 w2 Q / (dx0 dx1 dx2)."""
    X, V = np.asarray(X, dtype=np.float64).reshape(-1), np.asarray(V, dtype=np.float64).reshape(-1, depth)
    sh = None if Xshift is None else np.asarray(Xshift, dtype=np.float64).reshape(-1)
    shape = tuple(iupper[d] - ilower[d] + 1 + 2 * nugc[d] for d in range(ndim))[::-1]
    q = u.reshape((depth,) + shape)
    vol = dx[0] * dx[1]
    if ndim == 3:
        vol = vol * dx[2]
    for l, s in enumerate(np.asarray(indices).reshape(-1)):
        lo, hi, w = _user_stencils(kernel_fcn, stencil, ndim, dx, x_lower, ilower, iupper, nugc, X, sh, l, int(s))
        for d in range(depth):
            for i2 in range(lo[2], hi[2] + 1) if ndim == 3 else [0]:
                for i1 in range(lo[1], hi[1] + 1):
                    for i0 in range(lo[0], hi[0] + 1):
                        j = (i0 - ilower[0] + nugc[0], i1 - ilower[1] + nugc[1]) + ((i2 - ilower[2] + nugc[2],) if ndim == 3 else ())
                        ww = w[0][i0 - lo[0]] * w[1][i1 - lo[1]]
                        if ndim == 3:
                            ww = ww * w[2][i2 - lo[2]]
                        q[(d,) + j[::-1]] += ww * V[int(s), d] / vol
    return u


def get_cell_index(X, x_lower, x_upper, dx, ilower, iupper):
    X = _f64(X)
    ndim = len(dx)
    n = X.size // ndim
    out = np.zeros((n, ndim), dtype=np.int32)
    lib().le_oracle_get_cell_index(ndim, n, _dp(X), _dp(_f64(x_lower)), _dp(_f64(x_upper)), _dp(_f64(dx)),
                                   _ip(_i32(ilower)), _ip(_i32(iupper)), _ip(out))
    return out


def wrap_positions(X, x_lower, x_upper, periodic):
    X = _f64(X).copy()
    ndim = len(x_lower)
    esc = lib().le_oracle_wrap_positions(ndim, X.size // ndim, _dp(X), _dp(_f64(x_lower)), _dp(_f64(x_upper)),
                                         _ip(_i32(periodic)))
    return X, esc


# ----------------------------------------------------------------------------------------------
# patch geometry stand-in
# ----------------------------------------------------------------------------------------------
@dataclass
class PatchGeom:
    """What LEInteractor reads from SAMRAI's Patch/CartesianPatchGeometry."""
    lower: tuple  # patch box lower cell index
    upper: tuple  # patch box upper cell index (inclusive)
    x_lower: tuple
    x_upper: tuple
    dx: tuple
    gcw: tuple = ()

    @property
    def ndim(self):
        return len(self.lower)

    def side_shape(self, axis, gcw=None):
        g = self.gcw if gcw is None else gcw
        n = [self.upper[d] - self.lower[d] + 1 + (1 if d == axis else 0) + 2 * g[d] for d in range(self.ndim)]
        return tuple(reversed(n))

    def cell_shape(self, gcw=None):
        g = self.gcw if gcw is None else gcw
        n = [self.upper[d] - self.lower[d] + 1 + 2 * g[d] for d in range(self.ndim)]
        return tuple(reversed(n))

    def side_coords(self, axis, gcw=None):
        """Physical coordinates of every side of `axis` (incl. ghosts), as ndim arrays."""
        g = self.gcw if gcw is None else gcw
        ax = []
        for d in range(self.ndim):
            n = self.upper[d] - self.lower[d] + 1 + (1 if d == axis else 0) + 2 * g[d]
            i = np.arange(n) - g[d]
            ax.append(self.x_lower[d] + self.dx[d] * (i + (0.0 if d == axis else 0.5)))
        return np.meshgrid(*reversed(ax), indexing="ij")[::-1]

    def cell_coords(self, gcw=None):
        g = self.gcw if gcw is None else gcw
        ax = []
        for d in range(self.ndim):
            n = self.upper[d] - self.lower[d] + 1 + 2 * g[d]
            i = np.arange(n) - g[d]
            ax.append(self.x_lower[d] + self.dx[d] * (i + 0.5))
        return np.meshgrid(*reversed(ax), indexing="ij")[::-1]


def indices_in_box(X, pg: PatchGeom, box_lower, box_upper):
    """LEInteractor::buildLocalIndices, position form (LEInteractor.cpp:6088-6126)."""
    X = _f64(X)
    n = X.size // pg.ndim
    out = np.zeros(max(n, 1), dtype=np.int32)
    cnt = lib().le_oracle_indices_in_box(pg.ndim, n, _dp(X), _dp(_f64(pg.x_lower)), _dp(_f64(pg.x_upper)),
                                         _dp(_f64(pg.dx)), _ip(_i32(pg.lower)), _ip(_i32(pg.upper)),
                                         _ip(_i32(box_lower)), _ip(_i32(box_upper)), _ip(out))
    return out[:cnt].copy()


def _ptr_array(arrs):
    P = (C.POINTER(C.c_double) * len(arrs))()
    for i, a in enumerate(arrs):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        P[i] = _dp(a)
    return P


def side_interp(kernel, pg: PatchGeom, u_sides, X, indices, shifts=None, Q=None):
    """SideData interpolate with an explicit index list (LEInteractor.cpp:2402-2489)."""
    ndim = pg.ndim
    X = _f64(X)
    indices = _i32(indices)
    shifts = np.zeros(indices.size * ndim) if shifts is None else _f64(shifts)
    if Q is None:
        Q = np.zeros((X.size // ndim, ndim))
    lib().le_oracle_side_interp(KERNELS[kernel], ndim, _dp(_f64(pg.x_lower)), _dp(_f64(pg.dx)), _ip(_i32(pg.lower)),
                                _ip(_i32(pg.upper)), _ip(_i32(pg.gcw)), _ptr_array(u_sides), _ip(indices), _dp(shifts),
                                int(indices.size), _dp(X), _dp(Q))
    return Q


def side_spread(kernel, pg: PatchGeom, u_sides, X, Q, indices, shifts=None):
    """SideData spread with an explicit index list (LEInteractor.cpp:3627-3714)."""
    ndim = pg.ndim
    X = _f64(X)
    Q = _f64(Q)
    indices = _i32(indices)
    shifts = np.zeros(indices.size * ndim) if shifts is None else _f64(shifts)
    lib().le_oracle_side_spread(KERNELS[kernel], ndim, _dp(_f64(pg.x_lower)), _dp(_f64(pg.dx)), _ip(_i32(pg.lower)),
                                _ip(_i32(pg.upper)), _ip(_i32(pg.gcw)), _ptr_array(u_sides), _ip(indices), _dp(shifts),
                                int(indices.size), _dp(X), _dp(Q))
    return u_sides


def side_interp_positions(kernel, pg, u_sides, X, box_lower=None, box_upper=None):
    """Position-only SideData interpolate (LEInteractor.cpp:3045-3127)."""
    idx = indices_in_box(X, pg, pg.lower if box_lower is None else box_lower, pg.upper if box_upper is None else box_upper)
    return side_interp(kernel, pg, u_sides, X, idx)


def side_spread_positions(kernel, pg, u_sides, X, Q, box_lower=None, box_upper=None):
    """Position-only SideData spread (LEInteractor.cpp:4188-4263)."""
    idx = indices_in_box(X, pg, pg.lower if box_lower is None else box_lower, pg.upper if box_upper is None else box_upper)
    return side_spread(kernel, pg, u_sides, X, Q, idx)


def cell_interp_positions(kernel, pg, u_cell, depth, X, box_lower=None, box_upper=None):
    """Position-only CellData interpolate (LEInteractor.cpp:2805-2863); u_cell shape (depth, [n2,] n1, n0)."""
    idx = indices_in_box(X, pg, pg.lower if box_lower is None else box_lower, pg.upper if box_upper is None else box_upper)
    ndim = pg.ndim
    V = np.full((np.asarray(X).size // ndim, depth), np.finfo(np.float64).max)
    return interp_raw(kernel, ndim, pg.dx, pg.x_lower, depth, pg.lower, pg.upper, pg.gcw, u_cell, idx,
                      np.zeros(idx.size * ndim), X, V)


def cell_spread_positions(kernel, pg, u_cell, depth, X, Q, box_lower=None, box_upper=None):
    """Position-only CellData spread (LEInteractor.cpp:3951-4009)."""
    idx = indices_in_box(X, pg, pg.lower if box_lower is None else box_lower, pg.upper if box_upper is None else box_upper)
    ndim = pg.ndim
    return spread_raw(kernel, ndim, pg.dx, pg.x_lower, depth, idx, np.zeros(idx.size * ndim), X, Q, pg.lower, pg.upper,
                      pg.gcw, u_cell)


def _shifted_geom(pg, shifted):
    """Array geometry the private funnel receives for a centering (LEInteractor.cpp:3021-3026 node, :3313-3323 edge):
    x_lower - dx/2 and one more index in every shifted dimension."""
    xl = [pg.x_lower[d] - (0.5 * pg.dx[d] if shifted[d] else 0.0) for d in range(pg.ndim)]
    up = [pg.upper[d] + (1 if shifted[d] else 0) for d in range(pg.ndim)]
    return xl, up


def node_interp_positions(kernel, pg, u_node, depth, X):
    """Position-only NodeData interpolate (LEInteractor.cpp:2983-3043); u_node shape (depth, [n2+1,] n1+1, n0+1)."""
    idx = indices_in_box(X, pg, pg.lower, pg.upper)
    ndim = pg.ndim
    xl, up = _shifted_geom(pg, [True] * ndim)
    V = np.full((np.asarray(X).size // ndim, depth), np.finfo(np.float64).max)
    return interp_raw(kernel, ndim, pg.dx, xl, depth, pg.lower, up, pg.gcw, u_node, idx, np.zeros(idx.size * ndim), X, V)


def node_spread_positions(kernel, pg, u_node, depth, X, Q):
    """Position-only NodeData spread (LEInteractor.cpp:4122-4186)."""
    idx = indices_in_box(X, pg, pg.lower, pg.upper)
    ndim = pg.ndim
    xl, up = _shifted_geom(pg, [True] * ndim)
    return spread_raw(kernel, ndim, pg.dx, xl, depth, idx, np.zeros(idx.size * ndim), X, Q, pg.lower, up, pg.gcw, u_node)


def edge_interp_positions(kernel, pg, u_edges, X):
    """Position-only EdgeData interpolate (LEInteractor.cpp:3260-3340): per axis a scalar interpolation on the array
    shifted in every dimension but the axis; listed markers get component `axis` of Q."""
    idx = indices_in_box(X, pg, pg.lower, pg.upper)
    ndim = pg.ndim
    n = np.asarray(X).size // ndim
    Q = np.full((n, ndim), np.finfo(np.float64).max)
    for axis in range(ndim):
        xl, up = _shifted_geom(pg, [d != axis for d in range(ndim)])
        V = np.full((n, 1), np.finfo(np.float64).max)
        interp_raw(kernel, ndim, pg.dx, xl, 1, pg.lower, up, pg.gcw, u_edges[axis], idx, np.zeros(idx.size * ndim), X, V, axis=axis)
        Q[:, axis] = V[:, 0]
    return Q


def edge_spread_positions(kernel, pg, u_edges, X, Q):
    """Position-only EdgeData spread (LEInteractor.cpp:4386-4466)."""
    idx = indices_in_box(X, pg, pg.lower, pg.upper)
    ndim = pg.ndim
    Q = _f64(Q).reshape(-1, ndim)
    for axis in range(ndim):
        xl, up = _shifted_geom(pg, [d != axis for d in range(ndim)])
        spread_raw(kernel, ndim, pg.dx, xl, 1, idx, np.zeros(idx.size * ndim), X, np.ascontiguousarray(Q[:, axis:axis + 1]), pg.lower,
                   up, pg.gcw, u_edges[axis], axis=axis)
    return u_edges


# ----------------------------------------------------------------------------------------------
# level = set of patches on one refinement level of a Cartesian domain
# ----------------------------------------------------------------------------------------------
@dataclass
class Level:
    ndim: int
    domain_lower: tuple  # level cell index of the domain's lower corner (normally 0)
    domain_ncells: tuple  # cells of the (refined) physical domain per dimension
    x_lower: tuple
    x_upper: tuple
    periodic: tuple
    boxes: list  # [(lower tuple, upper tuple)] patch boxes on this level
    gcw: tuple
    dx: tuple = field(init=False)

    def __post_init__(self):
        self.dx = tuple((self.x_upper[d] - self.x_lower[d]) / self.domain_ncells[d] for d in range(self.ndim))

    def patch_geom(self, p) -> PatchGeom:
        lo, hi = self.boxes[p]
        # SAMRAI computes patch x_lower/x_upper as x_lo + dx * (index - domain_lower)
        xl = tuple(self.x_lower[d] + self.dx[d] * (lo[d] - self.domain_lower[d]) for d in range(self.ndim))
        xu = tuple(self.x_lower[d] + self.dx[d] * (hi[d] + 1 - self.domain_lower[d]) for d in range(self.ndim))
        return PatchGeom(tuple(lo), tuple(hi), xl, xu, self.dx, tuple(self.gcw))

    def domain_upper(self):
        return tuple(self.domain_lower[d] + self.domain_ncells[d] - 1 for d in range(self.ndim))


def bin_level(level: Level, X):
    """Marker -> (cell, owner patch) and the per-patch index lists.

    cell   = getCellIndex(X, grid_geom, ratio)        LDataManager.cpp:1475
    owner  = the patch whose box contains the cell    LDataManager.cpp:1476
    lists  = LIndexSetData::cacheLocalIndices         LIndexSetData.cpp:53-141
    Returns dict(cells, owner, patches=[dict(all_idx, all_shift, interior_mask)]).
    """
    ndim = level.ndim
    X = _f64(X)
    n = X.size // ndim
    cells = get_cell_index(X, level.x_lower, level.x_upper, level.dx, level.domain_lower, level.domain_upper())
    owner = np.full(n, -1, dtype=np.int32)
    for p, (lo, hi) in enumerate(level.boxes):
        m = np.ones(n, dtype=bool)
        for d in range(ndim):
            m &= (cells[:, d] >= lo[d]) & (cells[:, d] <= hi[d])
        owner[m] = p
    patches = []
    for p, (lo, hi) in enumerate(level.boxes):
        args = (ndim, n, _ip(cells), _ip(_i32(lo)), _ip(_i32(hi)), _ip(_i32(level.gcw)), _ip(_i32(level.domain_lower)),
                _ip(_i32(level.domain_ncells)), _ip(_i32(level.periodic)), _dp(_f64(level.dx)))
        cnt = lib().le_oracle_patch_lists(*args, None, None, None)
        idx = np.zeros(max(cnt, 1), dtype=np.int32)
        sh = np.zeros(max(cnt, 1) * ndim)
        interior = np.zeros(max(cnt, 1), dtype=np.int32)
        lib().le_oracle_patch_lists(*args, _ip(idx), _dp(sh), _ip(interior))
        patches.append(dict(all_idx=idx[:cnt].copy(), all_shift=sh[:cnt * ndim].copy(),
                            interior_mask=interior[:cnt].astype(bool)))
    return dict(cells=cells, owner=owner, patches=patches)


def fold_walls(level: Level, p, arrays, acoef, bcoef):
    """CartSideRobinPhysBdryOp::accumulateFromPhysicalBoundaryData for patch p, co-dimension one, LINEAR, homogeneous
    (ibtk/src/boundary/physical_boundary/CartSideRobinPhysBdryOp.cpp:552-617): the adjoint of the ghost-cell extrapolation,
    restated from fortran/cartphysbdryop3d.f.m4 (scrobinphysbdryop1x3d :787-905 for the component normal to the wall,
    ccrobinphysbdryop1x3d :78-168 for the transverse ones; adjoint_op = 1).  arrays[axis]: the patch's side arrays with
    ghosts, modified in place over the patch's own (side-box) extent in the transverse directions.
    acoef / bcoef [ndim][2][ndim components]."""
    ndim = level.ndim
    lo, hi = level.boxes[p]
    g = level.gcw
    acoef = np.asarray(acoef, dtype=np.float64).reshape(ndim, 2, ndim)
    bcoef = np.asarray(bcoef, dtype=np.float64).reshape(ndim, 2, ndim)
    for d in range(ndim):
        if level.periodic[d]:
            continue
        for side in (0, 1):
            touches = lo[d] == level.domain_lower[d] if side == 0 else hi[d] == level.domain_upper()[d]
            if not touches:
                continue
            ncell = hi[d] - lo[d] + 1
            h = level.dx[d]
            sgn = -1 if side == 0 else 1
            for comp in range(ndim):
                u = arrays[comp]
                a, b = acoef[d, side, comp], bcoef[d, side, comp]
                # view with the wall normal as the LAST python axis removed: index helper along array axis (ndim - 1 - d)
                ax = ndim - 1 - d
                sl = [slice(None)] * ndim
                for e in range(ndim):  # transverse extent: the component's side box (interior)
                    if e == d:
                        continue
                    n_int = hi[e] - lo[e] + 1 + (1 if e == comp else 0)
                    sl[ndim - 1 - e] = slice(g[e], g[e] + n_int)

                def at(i):
                    t = list(sl)
                    t[ax] = i
                    return tuple(t)
                if comp == d:
                    ib = g[d] if side == 0 else g[d] + ncell
                    dirichlet = abs(b) < 1e-12
                    if dirichlet:
                        u[at(ib)] = 0.0  # u_b = g / a with g = 0 (f.m4:864-865)
                    for i in range(1, g[d] + 1):
                        ug = u[at(ib + sgn * i)].copy()
                        fi = -1.0 if dirichlet else 1.0
                        fb = 2.0 if dirichlet else -a * (2.0 * i) * h / b
                        u[at(ib - sgn * i)] += fi * ug
                        u[at(ib)] += fb * ug
                else:
                    ii = g[d] if side == 0 else g[d] + ncell - 1
                    for i in range(g[d]):
                        nn = 1.0 + 2.0 * i
                        fi = -(a * nn * h - 2.0 * b) / (a * nn * h + 2.0 * b)
                        u[at(ii - sgn * i)] += fi * u[at(ii + sgn * (1 + i))]
    return arrays


def _amr_index(level: Level, p, axis, d):
    """index of element 0 of patch p's side array (component axis) along dimension d"""
    return level.boxes[p][0][d] - level.gcw[d]


def amr_refine_side(coarse: Level, fine: Level, ratio, c_arrays, f_arrays):
    """f_arrays[pf][axis] := CONSERVATIVE_LINEAR_REFINE(c_arrays[pc][axis]) (f_prolongation_scheds[ln]->fillData,
    LDataManager.cpp:611-614; registered src/IB/IBHierarchyIntegrator.cpp:374-377).  The operator is SAMRAI's
    CartesianSideDoubleConservativeLinearRefine (third party, IBSAMRAI2, not under /root/reference): restated from its
    published algorithm (Fortran cartclinrefsidedoub{2,3}d{0,1,2}) -- PARITY UNPINNED, no reference fixture holds its output.
    Every fine point (ghosts included) whose coarse stencil lies inside a coarse patch's array is written; coarse patches
    in list order.  Arrays are [z][y][x] with ghosts.  Returns the number of points written."""
    ndim = fine.ndim
    written = 0
    for pf in range(len(fine.boxes)):
        for pc in range(len(coarse.boxes)):
            for a in range(ndim):
                C_, F_ = c_arrays[pc][a], f_arrays[pf][a]
                idx_f, idx_c, ir = [], [], []
                empty = False
                for d in range(ndim):
                    side = 1 if d == a else 0
                    r = ratio[d]
                    cu_lo = coarse.boxes[pc][0][d] - coarse.gcw[d] + 1
                    cu_hi = coarse.boxes[pc][1][d] + coarse.gcw[d] + side - 1
                    flo = max(fine.boxes[pf][0][d] - fine.gcw[d], cu_lo * r)
                    fhi = min(fine.boxes[pf][1][d] + fine.gcw[d] + side, cu_hi * r + r - 1)
                    if fhi < flo:
                        empty = True
                        break
                    i = np.arange(flo, fhi + 1)
                    ic = np.floor_divide(i, r)
                    idx_f.append(i - _amr_index(fine, pf, a, d))
                    idx_c.append(ic - _amr_index(coarse, pc, a, d))
                    ir.append((i - ic * r).astype(np.float64))
                if empty:
                    continue

                def take(shift_d=None, s=0):
                    ix = [idx_c[d] + (s if d == shift_d else 0) for d in range(ndim)]
                    return C_[np.ix_(*ix[::-1])]
                c = take()
                v = c.copy()
                for d in range(ndim):
                    dm = c - take(d, -1)
                    dp = take(d, +1) - c
                    coef2 = 0.5 * (dm + dp)
                    bound = 2.0 * np.minimum(np.abs(dm), np.abs(dp))
                    slope = np.where(dm * dp > 0.0, np.copysign(np.minimum(np.abs(coef2), bound), coef2) / coarse.dx[d], 0.0)
                    if d == a:
                        delta = ir[d] * fine.dx[d]
                    else:
                        delta = (ir[d] + 0.5) * fine.dx[d] - coarse.dx[d] * 0.5
                    shp = [1] * ndim
                    shp[ndim - 1 - d] = -1
                    v = v + slope * delta.reshape(shp)
                F_[np.ix_(*idx_f[::-1])] = v
                written += v.size
    return written


def amr_coarsen_side(coarse: Level, fine: Level, ratio, c_arrays, f_arrays):
    """c_arrays[pc][axis] := CONSERVATIVE_COARSEN(f_arrays[pf][axis]) on the coarse patches' own sides tiled by a fine patch's own
    sides (f_synch_scheds[ln]->coarsenData, LDataManager.cpp:728-734; registered IBHierarchyIntegrator.cpp:369-372).  SAMRAI's
    CartesianSideDoubleWeightedAverage (third party; Fortran cartwgtavgsidedoub{2,3}d{0,1,2}), restated from its published
    algorithm: sum of fine * dAf over the fine sides of the coarse side (highest dimension outermost), divided by dAc --
    PARITY UNPINNED.  Returns the number of points written."""
    ndim = fine.ndim
    written = 0
    for pc in range(len(coarse.boxes)):
        for pf in range(len(fine.boxes)):
            for a in range(ndim):
                C_, F_ = c_arrays[pc][a], f_arrays[pf][a]
                ics, empty = [], False
                for d in range(ndim):
                    r = ratio[d]
                    flo, fhi = fine.boxes[pf][0][d], fine.boxes[pf][1][d]
                    lo = -((-flo) // r)
                    if d == a:
                        hi = (fhi + 1) // r
                        lo, hi = max(lo, coarse.boxes[pc][0][d]), min(hi, coarse.boxes[pc][1][d] + 1)
                    else:
                        hi = (fhi + 1) // r - 1
                        lo, hi = max(lo, coarse.boxes[pc][0][d]), min(hi, coarse.boxes[pc][1][d])
                    if hi < lo:
                        empty = True
                        break
                    ics.append(np.arange(lo, hi + 1))
                if empty:
                    continue
                dAf = dAc = 1.0
                for d in range(ndim):
                    if d != a:
                        dAf *= fine.dx[d]
                        dAc *= coarse.dx[d]
                rr = [1 if d == a else ratio[d] for d in range(ndim)] + [1] * (3 - ndim)
                s = np.zeros(tuple(len(ics[d]) for d in range(ndim))[::-1])
                for k2 in range(rr[2]):
                    for k1 in range(rr[1]):
                        for k0 in range(rr[0]):
                            k = (k0, k1, k2)
                            ix = [ics[d] * ratio[d] + (0 if d == a else k[d]) - _amr_index(fine, pf, a, d) for d in range(ndim)]
                            s = s + F_[np.ix_(*ix[::-1])] * dAf
                ixc = [ics[d] - _amr_index(coarse, pc, a, d) for d in range(ndim)]
                C_[np.ix_(*ixc[::-1])] = s / dAc
                written += s.size
    return written


def sc_interp_op(level: Level, X, dof_index, interp_fcn="IB_4"):
    """PETScMatUtilities::constructPatchLevelSCInterpOp (ibtk/src/math/PETScMatUtilities.cpp:783-1020) as (cols, vals), each
    [ndim * n, stencil^ndim], row ndim * k + axis, entries in box-iterator order (x fastest).  interp_fcn: ib_4_interp_fcn /
    pwl_interp_fcn of PETScMatUtilities.h:156-176.  dof_index[p][axis]: int arrays [z][y][x] with the level's ghost width."""
    ndim = level.ndim
    X = np.asarray(X, dtype=np.float64).reshape(-1, ndim)
    n = X.shape[0]
    S = 4 if interp_fcn == "IB_4" else 2
    dom_lo, dom_hi = level.domain_lower, level.domain_upper()
    cells = get_cell_index(X, level.x_lower, level.x_upper, level.dx, dom_lo, dom_hi).reshape(-1, ndim)
    cols = np.full((ndim * n, S ** ndim), -1, dtype=np.int32)
    vals = np.zeros((ndim * n, S ** ndim))

    def weights(r):
        if S == 4:
            q = np.sqrt(-7.0 + 12.0 * r - 4.0 * r * r)
            return np.stack([0.125 * (5.0 - 2.0 * r - q), 0.125 * (5.0 - 2.0 * r + q), 0.125 * (-1.0 + 2.0 * r + q),
                             0.125 * (-1.0 + 2.0 * r - q)], axis=1)
        return np.stack([1.0 - r, r], axis=1)
    # the patch of every marker: interior first, then the first ghost layer (:861-877), patches in list order
    pnum = np.full(n, -1)
    for growth in (0, 1):
        for p, (lo, hi) in enumerate(level.boxes):
            inside = np.all((cells >= np.array(lo) - growth) & (cells <= np.array(hi) + growth), axis=1)
            pnum[(pnum < 0) & inside] = p
    x_cell = np.stack([(cells[:, d] - dom_lo[d] + 0.5) * level.dx[d] + level.x_lower[d] for d in range(ndim)], axis=1)
    for axis in range(ndim):
        lower = np.zeros((n, ndim), dtype=np.int64)
        w = []
        for d in range(ndim):
            if d == axis:
                lower[:, d] = cells[:, d] - S // 2 + 1
            else:
                lower[:, d] = np.where(X[:, d] <= x_cell[:, d], cells[:, d] - S // 2, cells[:, d] - S // 2 + 1)
            x_sl = ((lower[:, d] - dom_lo[d]).astype(np.float64) + (0.0 if d == axis else 0.5)) * level.dx[d] + level.x_lower[d]
            w.append(weights((X[:, d] - x_sl) / level.dx[d]))
        for p, (lo, hi) in enumerate(level.boxes):
            m = np.nonzero(pnum == p)[0]
            if m.size == 0:
                continue
            dof = np.asarray(dof_index[p][axis])
            e = 0
            for kz in range(S if ndim == 3 else 1):
                for ky in range(S):
                    for kx in range(S):
                        k = (kx, ky, kz)
                        v = w[0][m, kx] * w[1][m, ky]
                        if ndim == 3:
                            v = v * w[2][m, kz]
                        j = [lower[m, d] + k[d] - (lo[d] - level.gcw[d]) for d in range(ndim)]
                        cols[ndim * m + axis, e] = dof[tuple(j[::-1])]
                        vals[ndim * m + axis, e] = v
                        e += 1
    return cols, vals


def ghost_accumulate(level: Level, arrays, centering="side"):
    """SAMRAIGhostDataAccumulator::accumulateGhostData on one level.

    arrays[p][axis] (side) or arrays[p][0] (cell): per-patch arrays including ghosts.  Every copy
    of a DOF (interior copies of shared faces and all ghost copies, periodic images included) is
    summed into one value, which is then written back to every copy
    (SAMRAIGhostDataAccumulator.cpp:327-344).  Copies with no owner on this level (outside a
    non-periodic domain, or under no patch) keep their values (DOF index -1, :114,231).
    Modifies arrays in place.
    """
    ndim = level.ndim
    ncomp = ndim if centering == "side" else 1
    for axis in range(ncomp):
        # global DOF space of this component
        ext = [level.domain_ncells[d] + (1 if (centering == "side" and d == axis and not level.periodic[d]) else 0)
               for d in range(ndim)]
        total = np.zeros(tuple(reversed(ext)))
        owned = np.zeros(tuple(reversed(ext)), dtype=bool)
        maps = []
        for p, (lo, hi) in enumerate(level.boxes):
            a = arrays[p][axis]
            gi = []
            valid = []
            for d in range(ndim):
                n = hi[d] - lo[d] + 1 + (1 if (centering == "side" and d == axis) else 0) + 2 * level.gcw[d]
                g = np.arange(n) + lo[d] - level.gcw[d] - level.domain_lower[d]
                if level.periodic[d]:
                    v = np.ones(n, dtype=bool)
                    g = np.mod(g, level.domain_ncells[d])
                else:
                    v = (g >= 0) & (g < ext[d])
                    g = np.clip(g, 0, ext[d] - 1)
                gi.append(g)
                valid.append(v)
                # interior (owned) range of this patch
            mesh = np.meshgrid(*reversed(gi), indexing="ij")
            vmesh = np.meshgrid(*reversed(valid), indexing="ij")
            vm = np.logical_and.reduce(vmesh)
            maps.append((mesh, vm))
            # mark DOFs owned by some patch interior
            sl = []
            for d in reversed(range(ndim)):
                n_int = hi[d] - lo[d] + 1 + (1 if (centering == "side" and d == axis) else 0)
                sl.append(slice(level.gcw[d], level.gcw[d] + n_int))
            im = np.zeros(a.shape, dtype=bool)
            im[tuple(sl)] = True
            owned[tuple(m[im & vm] for m in mesh)] = True
        for p in range(len(level.boxes)):
            mesh, vm = maps[p]
            a = arrays[p][axis]
            ok = vm & owned[tuple(mesh)]
            np.add.at(total, tuple(m[ok] for m in mesh), a[ok])
        for p in range(len(level.boxes)):
            mesh, vm = maps[p]
            a = arrays[p][axis]
            ok = vm & owned[tuple(mesh)]
            a[ok] = total[tuple(m[ok] for m in mesh)]
    return arrays


# ----------------------------------------------------------------------------------------------
# CPU baseline (reference parallel model on OpenMP threads); bench.py only
# ----------------------------------------------------------------------------------------------
class Baseline:
    def __init__(self, ndim, N, npatch, gcw, x_lower, x_upper, X, field_seed=0):
        self.ndim = ndim
        self.X = _f64(X)
        self.n = self.X.size // ndim
        self._h = lib().le_oracle_baseline_create(ndim, _ip(_i32(N)), _ip(_i32(npatch)), int(gcw), _dp(_f64(x_lower)),
                                                  _dp(_f64(x_upper)), self.n, _dp(self.X),
                                                  C.c_ulonglong(field_seed))
        self._h = C.c_void_p(self._h)

    def step(self, kernel, F, U):
        lib().le_oracle_baseline_step(self._h, KERNELS[kernel], _dp(self.X), _dp(F), _dp(U))

    def zero_f(self):
        lib().le_oracle_baseline_zero_f(self._h)

    def npatch(self):
        return int(lib().le_oracle_baseline_npatch(self._h))

    def patch_array(self, which, q, axis):
        """A copy of worker q's private array (`which` = "u" or "f", component `axis`), flat, x fastest, ghosts included."""
        fn = lib().le_oracle_baseline_f if which == "f" else lib().le_oracle_baseline_u
        n = int(lib().le_oracle_baseline_array_size(self._h, int(q), int(axis)))
        return np.ctypeslib.as_array(fn(self._h, int(q), int(axis)), shape=(n,)).copy()

    @staticmethod
    def threads():
        """All cores this process may run on (torchrun's OMP_NUM_THREADS=1 default is overridden)."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        lib().le_oracle_baseline_set_threads(int(n))
        return lib().le_oracle_baseline_threads()

    def close(self):
        if self._h:
            lib().le_oracle_baseline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# N1: Lagrangian forces (oracle/le_force.c) and N2: the structure files of IBStandardInitializer
# ------------------------------------------------------------------------------------------------
def lagrangian_force(ndim, X, U, springs=None, beams=None, targets=None):
    """IBMethod::computeLagrangianForce + IBStandardForceGen::computeLagrangianForce
    (IBMethod.cpp:834-858, IBStandardForceGen.cpp:253-303): F = 0, then springs, beams, target points.
    springs = (mastr, slave, kappa, rest); beams = (mastr, next, prev, rigidity, curvature[n][ndim]);
    targets = (idx, kappa, eta, X0[n][ndim]).  Node arrays are AoS [n][ndim]."""
    X, U = _f64(X), _f64(U)
    F = np.zeros_like(X)
    if springs is not None and len(springs[0]):
        m, s, k, r = springs
        lib().le_oracle_spring_force(ndim, len(m), _ip(_i32(m)), _ip(_i32(s)), _dp(_f64(k)), _dp(_f64(r)), _dp(X), _dp(F))
    if beams is not None and len(beams[0]):
        m, nx, pv, k, c = beams
        lib().le_oracle_beam_force(ndim, len(m), _ip(_i32(m)), _ip(_i32(nx)), _ip(_i32(pv)), _dp(_f64(k)), _dp(_f64(c)), _dp(X),
                                   _dp(F))
    if targets is not None and len(targets[0]):
        i, k, e, x0 = targets
        lib().le_oracle_target_force(ndim, len(i), _ip(_i32(i)), _dp(_f64(k)), _dp(_f64(e)), _dp(_f64(x0)), _dp(X), _dp(U), _dp(F))
    return F


def _discard_comments(line):
    """IBStandardInitializer.cpp:65-87: drop everything after '!', '#' or '%'."""
    for ch in "!#%":
        line = line.split(ch, 1)[0]
    return line


def _structure_lines(path):
    with open(path) as f:
        lines = f.read().split("\n")
    count = int(_discard_comments(lines[0]).split()[0])
    if count <= 0:
        raise ValueError("invalid count on line 1")
    rows = [_discard_comments(l).split() for l in lines[1:1 + count]]
    if len(rows) < count or any(len(r) == 0 for r in rows):
        raise ValueError("premature end of file")
    return count, rows


def read_vertex_file(path, ndim):
    """IBStandardInitializer::readVertexFiles (IBStandardInitializer.cpp:184-294); no shift / scale."""
    n, rows = _structure_lines(path)
    return np.array([[float(v) for v in r[:ndim]] for r in rows], dtype=np.float64).reshape(n, ndim)


def read_spring_file(path, n_vertices, offset=0):
    """readSpringFiles (IBStandardInitializer.cpp:297-528): 'mastr slave kappa rest [fcn_idx ...]'; the edge is
    stored with the smaller index first (:478-481), duplicates are skipped (:482-499)."""
    _, rows = _structure_lines(path)
    seen, out = set(), []
    for r in rows:
        a, b, k, rest = int(r[0]), int(r[1]), float(r[2]), float(r[3])
        if not (0 <= a < n_vertices and 0 <= b < n_vertices) or k < 0.0 or rest < 0.0:
            raise ValueError("invalid spring entry")
        fcn = int(r[4]) if len(r) > 4 else 0
        a, b = a + offset, b + offset
        if a > b:
            a, b = b, a
        if (a, b) in seen:
            continue
        seen.add((a, b))
        out.append((a, b, k, rest, fcn))
    return (np.array([o[0] for o in out], dtype=np.int32), np.array([o[1] for o in out], dtype=np.int32),
            np.array([o[2] for o in out]), np.array([o[3] for o in out]), np.array([o[4] for o in out], dtype=np.int32))


def read_beam_file(path, n_vertices, ndim, offset=0):
    """readBeamFiles (IBStandardInitializer.cpp:766-1002): 'prev curr next bend [curvature x ndim]'."""
    _, rows = _structure_lines(path)
    prev, curr, nxt, bend, curv = [], [], [], [], []
    seen = set()
    for r in rows:
        p, c, n, b = int(r[0]), int(r[1]), int(r[2]), float(r[3])
        if not all(0 <= v < n_vertices for v in (p, c, n)) or b < 0.0:
            raise ValueError("invalid beam entry")
        cv = [float(v) for v in r[4:4 + ndim]] if len(r) >= 4 + ndim else [0.0] * ndim
        key = (c + offset, n + offset, p + offset)
        if key in seen:
            continue
        seen.add(key)
        prev.append(p + offset), curr.append(c + offset), nxt.append(n + offset), bend.append(b), curv.append(cv)
    return (np.array(prev, dtype=np.int32), np.array(curr, dtype=np.int32), np.array(nxt, dtype=np.int32), np.array(bend),
            np.array(curv, dtype=np.float64).reshape(-1, ndim))


def read_target_file(path, n_vertices, offset=0):
    """readTargetPointFiles (IBStandardInitializer.cpp:1322-1517): 'idx kappa [eta]'."""
    _, rows = _structure_lines(path)
    idx, kappa, eta = [], [], []
    for r in rows:
        i, k = int(r[0]), float(r[1])
        e = float(r[2]) if len(r) > 2 else 0.0
        if not 0 <= i < n_vertices or k < 0.0 or e < 0.0:
            raise ValueError("invalid target point entry")
        idx.append(i + offset), kappa.append(k), eta.append(e)
    return np.array(idx, dtype=np.int32), np.array(kappa), np.array(eta)
