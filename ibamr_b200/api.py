"""Python mirror of the reference's operator interface for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference so that the parity tests read
like the reference's own tests:

  LEInteractor.interpolate / .spread / .getStencilSize / .getMinimumGhostWidth / .isKnownKernel
      ibtk/include/ibtk/LEInteractor.h:99-117, 566-575, 1132-1141, 184-192, 704-712
  IBMethodB200.spreadForce / .interpolateVelocity / .beginDataRedistribution / .endDataRedistribution
      include/ibamr/IBStrategy.h:276-280, 338-342, 455, 464 (as overridden by src/IB/IBMethod.cpp)

Box / Patch / SideData / CellData are minimal stand-ins for the SAMRAI types (SAMRAI itself is a
third-party dependency that is not in this image).  Everything computes on the GPU through
libibk.so; a failed call raises IBKError carrying ibk_last_error (the reference aborts through
TBOX_ERROR at the same places).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ArrayDesc, LevelDesc, PatchDesc


# marker columns (include/ibk.h IBK_COL_*): X is the one spread / interpolate / re-bin work on
COLUMNS = {"X": 0, "U": 1, "F": 2, "X_current": 3, "X_new": 4, "aux": 5}


class IBKError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ibk error {code}: {msg}")
        self.code = code


IBK_ERR_INVALID, IBK_ERR_CUDA, IBK_ERR_UNKNOWN_KERNEL, IBK_ERR_GHOST_WIDTH, IBK_ERR_DEPTH, IBK_ERR_STATE, IBK_ERR_ESCAPED = (
    -1, -2, -3, -4, -5, -6, -7)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One per (process, CUDA device)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ibk_ctx_create(device, C.byref(h))
        if rc != 0:
            raise IBKError(rc, "ibk_ctx_create failed (no CUDA device? ibamr_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise IBKError(rc, self.lib.ibk_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr: int):
        self.check(self.lib.ibk_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(self.lib.ibk_ctx_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.lib.ibk_ctx_launch_count(self.h))

    def enable_timing(self, on=True):
        self.check(self.lib.ibk_ctx_enable_timing(self.h, int(on)))

    def last_ms(self, which: int) -> float:
        ms = C.c_float()
        self.check(self.lib.ibk_ctx_last_ms(self.h, which, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ibk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX = None


def default_context() -> Context:
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context(0)
    return _DEFAULT_CTX


# ------------------------------------------------------------------------------------------------
# SAMRAI stand-ins
# ------------------------------------------------------------------------------------------------
@dataclass
class Box:
    lower: tuple
    upper: tuple

    def __eq__(self, other):
        return tuple(self.lower) == tuple(other.lower) and tuple(self.upper) == tuple(other.upper)


@dataclass
class Patch:
    """hier::Patch + geom::CartesianPatchGeometry: box, x_lower, x_upper, dx, boundary flags."""
    box: Box
    x_lower: tuple
    x_upper: tuple
    dx: tuple
    touches_regular_bdry: bool = False

    @property
    def ndim(self):
        return len(self.box.lower)

    def getBox(self):
        return self.box


class SideData:
    """pdat::SideData<NDIM,double>: one array per axis over toSideBox(box, axis) grown by gcw.
    Arrays are C-ordered numpy arrays of shape ([n2,] n1, n0) (= the Fortran memory order)."""

    def __init__(self, box: Box, depth: int, gcw, arrays=None):
        self.box, self.depth = box, depth
        ndim = len(box.lower)
        self.gcw = tuple(gcw) if np.ndim(gcw) else (int(gcw),) * ndim
        if arrays is None:
            arrays = [np.zeros(self.shape(axis)) for axis in range(ndim)]
        self.arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]

    def shape(self, axis):
        ndim = len(self.box.lower)
        n = [self.box.upper[d] - self.box.lower[d] + 1 + (1 if d == axis else 0) + 2 * self.gcw[d] for d in range(ndim)]
        return ((self.depth,) if self.depth > 1 else ()) + tuple(reversed(n))

    def getDepth(self):
        return self.depth

    def getGhostCellWidth(self):
        return self.gcw

    def getPointer(self, axis):
        return self.arrays[axis]

    def fillAll(self, v):
        for a in self.arrays:
            a[...] = v


class CellData:
    """pdat::CellData<NDIM,double>: array of shape (depth, [n2,] n1, n0)."""

    def __init__(self, box: Box, depth: int, gcw, array=None):
        self.box, self.depth = box, depth
        ndim = len(box.lower)
        self.gcw = tuple(gcw) if np.ndim(gcw) else (int(gcw),) * ndim
        n = [box.upper[d] - box.lower[d] + 1 + 2 * self.gcw[d] for d in range(ndim)]
        shape = (depth,) + tuple(reversed(n))
        self.array = np.zeros(shape) if array is None else np.ascontiguousarray(array, dtype=np.float64).reshape(shape)

    def getDepth(self):
        return self.depth

    def getGhostCellWidth(self):
        return self.gcw

    def getPointer(self):
        return self.array

    def fillAll(self, v):
        self.array[...] = v


class NodeData:
    """pdat::NodeData<NDIM,double>: array of shape (depth, [n2+1,] n1+1, n0+1) over toNodeBox(box) grown by gcw."""

    def __init__(self, box: Box, depth: int, gcw, array=None):
        self.box, self.depth = box, depth
        ndim = len(box.lower)
        self.gcw = tuple(gcw) if np.ndim(gcw) else (int(gcw),) * ndim
        n = [box.upper[d] - box.lower[d] + 2 + 2 * self.gcw[d] for d in range(ndim)]
        shape = (depth,) + tuple(reversed(n))
        self.array = np.zeros(shape) if array is None else np.ascontiguousarray(array, dtype=np.float64).reshape(shape)

    def getDepth(self):
        return self.depth

    def getGhostCellWidth(self):
        return self.gcw

    def getPointer(self):
        return self.array


class EdgeData:
    """pdat::EdgeData<NDIM,double>: one array per axis over toEdgeBox(box, axis) (one more point in every dimension but
    the axis) grown by gcw."""

    def __init__(self, box: Box, depth: int, gcw, arrays=None):
        self.box, self.depth = box, depth
        ndim = len(box.lower)
        self.gcw = tuple(gcw) if np.ndim(gcw) else (int(gcw),) * ndim
        if arrays is None:
            arrays = [np.zeros(self.shape(axis)) for axis in range(ndim)]
        self.arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]

    def shape(self, axis):
        ndim = len(self.box.lower)
        n = [self.box.upper[d] - self.box.lower[d] + 1 + (0 if d == axis else 1) + 2 * self.gcw[d] for d in range(ndim)]
        return tuple(reversed(n))

    def getDepth(self):
        return self.depth

    def getGhostCellWidth(self):
        return self.gcw

    def getPointer(self, axis):
        return self.arrays[axis]


def _patch_desc(patch: Patch, gcw) -> PatchDesc:
    pd = PatchDesc()
    ndim = patch.ndim
    pd.ndim = ndim
    for d in range(ndim):
        pd.lower[d] = patch.box.lower[d]
        pd.upper[d] = patch.box.upper[d]
        pd.gcw[d] = gcw[d]
        pd.x_lower[d] = patch.x_lower[d]
        pd.x_upper[d] = patch.x_upper[d]
        pd.dx[d] = patch.dx[d]
    pd.touches_physical_bdry = int(patch.touches_regular_bdry)
    return pd


class LEInteractor:
    """Static interface of IBTK::LEInteractor for the in-scope kernels (LEInteractor.h:75)."""

    ctx = None

    @classmethod
    def _ctx(cls) -> Context:
        return cls.ctx or default_context()

    @staticmethod
    def isKnownKernel(kernel_fcn: str) -> bool:
        return bool(_lib.load().ibk_is_known_kernel(kernel_fcn.encode()))

    @staticmethod
    def getStencilSize(kernel_fcn: str) -> int:
        r = _lib.load().ibk_get_stencil_size(kernel_fcn.encode())
        if r < 0:
            raise IBKError(r, f"LEInteractor::getStencilSize(): Unknown kernel function {kernel_fcn}")
        return r

    @staticmethod
    def getMinimumGhostWidth(kernel_fcn: str) -> int:
        r = _lib.load().ibk_get_minimum_ghost_width(kernel_fcn.encode())
        if r < 0:
            raise IBKError(r, f"LEInteractor::getMinimumGhostWidth(): Unknown kernel function {kernel_fcn}")
        return r

    # -- position-only forms (LEInteractor.h:566-575 / 1132-1141 and the CellData twins) ---------
    @classmethod
    def interpolate(cls, Q_data, Q_depth, X_data, X_depth, q_data, patch: Patch, interp_box: Box, interp_fcn="IB_4"):
        ctx = cls._ctx()
        X = _f64(X_data).reshape(-1)
        assert Q_data.dtype == np.float64 and Q_data.flags.c_contiguous, "Q_data must be a contiguous float64 array"
        Q = Q_data.reshape(-1)
        pd = _patch_desc(patch, q_data.getGhostCellWidth())
        lo, hi = _i32(interp_box.lower), _i32(interp_box.upper)
        if isinstance(q_data, (SideData, EdgeData)):
            P = (C.POINTER(C.c_double) * patch.ndim)(*[_dp(a) for a in q_data.arrays])
            fn = ctx.lib.ibk_side_interpolate_host if isinstance(q_data, SideData) else ctx.lib.ibk_edge_interpolate_host
            rc = fn(ctx.h, interp_fcn.encode(), C.byref(pd), P, q_data.getDepth(), _ip(lo), _ip(hi), _dp(X), X.size, X_depth,
                    _dp(Q), Q.size, Q_depth)
        elif isinstance(q_data, NodeData):
            rc = ctx.lib.ibk_node_interpolate_host(ctx.h, interp_fcn.encode(), C.byref(pd), _dp(q_data.array),
                                                   q_data.getDepth(), _ip(lo), _ip(hi), _dp(X), X.size, X_depth, _dp(Q),
                                                   Q.size, Q_depth)
        else:
            rc = ctx.lib.ibk_cell_interpolate_host(ctx.h, interp_fcn.encode(), C.byref(pd), _dp(q_data.array),
                                                   q_data.getDepth(), _ip(lo), _ip(hi), _dp(X), X.size, X_depth, _dp(Q),
                                                   Q.size, Q_depth)
        ctx.check(rc)
        return Q_data

    @classmethod
    def spread(cls, q_data, Q_data, Q_depth, X_data, X_depth, patch: Patch, spread_box: Box, spread_fcn="IB_4"):
        ctx = cls._ctx()
        X = _f64(X_data).reshape(-1)
        Q = _f64(Q_data).reshape(-1)
        pd = _patch_desc(patch, q_data.getGhostCellWidth())
        lo, hi = _i32(spread_box.lower), _i32(spread_box.upper)
        if isinstance(q_data, (SideData, EdgeData)):
            P = (C.POINTER(C.c_double) * patch.ndim)(*[_dp(a) for a in q_data.arrays])
            fn = ctx.lib.ibk_side_spread_host if isinstance(q_data, SideData) else ctx.lib.ibk_edge_spread_host
            rc = fn(ctx.h, spread_fcn.encode(), C.byref(pd), P, q_data.getDepth(), _ip(lo), _ip(hi), _dp(X), X.size, X_depth,
                    _dp(Q), Q.size, Q_depth)
        elif isinstance(q_data, NodeData):
            rc = ctx.lib.ibk_node_spread_host(ctx.h, spread_fcn.encode(), C.byref(pd), _dp(q_data.array), q_data.getDepth(),
                                              _ip(lo), _ip(hi), _dp(X), X.size, X_depth, _dp(Q), Q.size, Q_depth)
        else:
            rc = ctx.lib.ibk_cell_spread_host(ctx.h, spread_fcn.encode(), C.byref(pd), _dp(q_data.array), q_data.getDepth(),
                                              _ip(lo), _ip(hi), _dp(X), X.size, X_depth, _dp(Q), Q.size, Q_depth)
        ctx.check(rc)
        return q_data

    # -- index-set forms (LEInteractor.h:184-192 / 704-712): the flat lists LIndexSetData caches ---
    @classmethod
    def interpolate_indexed(cls, Q_data, X_data, local_indices, periodic_shifts, q_data: SideData, patch: Patch,
                            interp_fcn="IB_4"):
        ctx = cls._ctx()
        X = _f64(X_data).reshape(-1)
        Q = Q_data.reshape(-1)
        idx = _i32(local_indices)
        sh = None if periodic_shifts is None else _f64(periodic_shifts).reshape(-1)
        pd = _patch_desc(patch, q_data.getGhostCellWidth())
        P = (C.POINTER(C.c_double) * patch.ndim)(*[_dp(a) for a in q_data.arrays])
        rc = ctx.lib.ibk_side_interpolate_indexed_host(ctx.h, interp_fcn.encode(), C.byref(pd), P, _ip(idx),
                                                       None if sh is None else _dp(sh), idx.size, _dp(X),
                                                       X.size // patch.ndim, _dp(Q))
        ctx.check(rc)
        return Q_data

    @classmethod
    def spread_indexed(cls, q_data: SideData, Q_data, X_data, local_indices, periodic_shifts, patch: Patch, spread_fcn="IB_4"):
        ctx = cls._ctx()
        X = _f64(X_data).reshape(-1)
        Q = _f64(Q_data).reshape(-1)
        idx = _i32(local_indices)
        sh = None if periodic_shifts is None else _f64(periodic_shifts).reshape(-1)
        pd = _patch_desc(patch, q_data.getGhostCellWidth())
        P = (C.POINTER(C.c_double) * patch.ndim)(*[_dp(a) for a in q_data.arrays])
        rc = ctx.lib.ibk_side_spread_indexed_host(ctx.h, spread_fcn.encode(), C.byref(pd), P, _ip(idx),
                                                  None if sh is None else _dp(sh), idx.size, _dp(X), X.size // patch.ndim,
                                                  _dp(Q))
        ctx.check(rc)
        return q_data


# ------------------------------------------------------------------------------------------------
# raw funnel (seam B4)
# ------------------------------------------------------------------------------------------------
def _array_desc(ndim, depth, dx, x_lower, x_upper, ilower, iupper, nugc) -> ArrayDesc:
    ad = ArrayDesc()
    ad.ndim, ad.depth = ndim, depth
    for d in range(ndim):
        ad.dx[d], ad.x_lower[d], ad.x_upper[d] = dx[d], x_lower[d], x_upper[d]
        ad.ilower[d], ad.iupper[d], ad.nugc[d] = ilower[d], iupper[d], nugc[d]
    return ad


def raw_interp_host(kernel: str, ndim, dx, x_lower, x_upper, depth, ilower, iupper, nugc, u, indices, Xshift, X, V, ctx=None):
    """lagrangian_<kernel>_interp{2,3}d with host arrays (3d.f.m4:1203-1209 argument meaning)."""
    ctx = ctx or default_context()
    ad = _array_desc(ndim, depth, dx, x_lower, x_upper, ilower, iupper, nugc)
    k = ctx.lib.ibk_kernel_from_string(kernel.encode())
    u, X = _f64(u), _f64(X).reshape(-1)
    idx = _i32(indices)
    sh = None if Xshift is None else _f64(Xshift).reshape(-1)
    assert V.dtype == np.float64 and V.flags.c_contiguous
    rc = ctx.lib.ibk_raw_interp_host(ctx.h, k, C.byref(ad), _dp(u), _ip(idx), None if sh is None else _dp(sh), idx.size,
                                     _dp(X), X.size // ndim, _dp(V.reshape(-1)))
    ctx.check(rc)
    return V


def raw_spread_host(kernel: str, ndim, dx, x_lower, x_upper, depth, indices, Xshift, X, V, ilower, iupper, nugc, u, ctx=None):
    """lagrangian_<kernel>_spread{2,3}d with host arrays (3d.f.m4:1344-1350 argument meaning)."""
    ctx = ctx or default_context()
    ad = _array_desc(ndim, depth, dx, x_lower, x_upper, ilower, iupper, nugc)
    k = ctx.lib.ibk_kernel_from_string(kernel.encode())
    X, V = _f64(X).reshape(-1), _f64(V).reshape(-1)
    idx = _i32(indices)
    sh = None if Xshift is None else _f64(Xshift).reshape(-1)
    assert u.dtype == np.float64 and u.flags.c_contiguous
    rc = ctx.lib.ibk_raw_spread_host(ctx.h, k, C.byref(ad), _ip(idx), None if sh is None else _dp(sh), idx.size, _dp(X),
                                     X.size // ndim, _dp(V), _dp(u.reshape(-1)))
    ctx.check(rc)
    return u


# ------------------------------------------------------------------------------------------------
# device-resident level: IBMethod / LDataManager roles (seams B1/B2)
# ------------------------------------------------------------------------------------------------
class IBMethodB200:
    """The IBStrategy-shaped object an integrator would hold (include/ibamr/IBStrategy.h).

    The level (patch boxes of THIS process, domain, periodicity, ghost width) replaces the
    PatchHierarchy argument; u / f live on the device (grid_upload / grid_download move SAMRAI
    SideData arrays in and out); marker columns X, U, F live on the device in binned order and are
    exchanged with the host in Lagrangian index order (LData AoS layout)."""

    def __init__(self, ndim, domain_lower, domain_upper, x_lower, x_upper, periodic, boxes, gcw=None, kernel_fcn="IB_4",
                 ctx: Context | None = None, error_if_points_leave_domain=False):
        self.ctx = ctx or default_context()
        self.ndim = ndim
        self.kernel_fcn = kernel_fcn  # IBMethod::d_interp_kernel_fcn / d_spread_kernel_fcn default "IB_4" (IBMethod.h:602)
        self.interp_kernel_fcn = kernel_fcn
        self.spread_kernel_fcn = kernel_fcn
        self.error_if_points_leave_domain = error_if_points_leave_domain
        g = LEInteractor.getMinimumGhostWidth(kernel_fcn) if gcw is None else gcw
        self.gcw = tuple(g) if np.ndim(g) else (int(g),) * ndim
        self.boxes = [(tuple(lo), tuple(hi)) for lo, hi in boxes]
        ld = LevelDesc()
        ld.ndim, ld.n_patches = ndim, len(self.boxes)
        for d in range(ndim):
            ld.domain_lower[d], ld.domain_upper[d] = domain_lower[d], domain_upper[d]
            ld.x_lower[d], ld.x_upper[d] = x_lower[d], x_upper[d]
            ld.periodic[d], ld.gcw[d] = int(periodic[d]), self.gcw[d]
        self._plo = _i32([b[0] for b in self.boxes]).reshape(-1) if self.boxes else _i32([0])
        self._phi = _i32([b[1] for b in self.boxes]).reshape(-1) if self.boxes else _i32([0])
        ld.patch_lower, ld.patch_upper = _ip(self._plo), _ip(self._phi)
        self.ctx.check(self.ctx.lib.ibk_level_create(self.ctx.h, C.byref(ld)))
        self.n_markers = 0

    # IBStrategy::getMinimumGhostCellWidth (IBMethod.cpp:266-270)
    def getMinimumGhostCellWidth(self):
        return max(LEInteractor.getMinimumGhostWidth(self.interp_kernel_fcn),
                   LEInteractor.getMinimumGhostWidth(self.spread_kernel_fcn))

    def side_shape(self, patch, axis):
        lo, hi = self.boxes[patch]
        n = [hi[d] - lo[d] + 1 + (1 if d == axis else 0) + 2 * self.gcw[d] for d in range(self.ndim)]
        return tuple(reversed(n))

    # -- Eulerian data -----------------------------------------------------------------------------
    def grid_upload(self, which, patch, axis, array):
        a = _f64(array)
        assert a.shape == self.side_shape(patch, axis), (a.shape, self.side_shape(patch, axis))
        self.ctx.check(self.ctx.lib.ibk_grid_upload(self.ctx.h, {"u": 0, "f": 1}[which], patch, axis, _dp(a)))

    def grid_download(self, which, patch, axis):
        a = np.zeros(self.side_shape(patch, axis))
        self.ctx.check(self.ctx.lib.ibk_grid_download(self.ctx.h, {"u": 0, "f": 1}[which], patch, axis, _dp(a)))
        return a

    def grid_upload_async(self, which, patch, axis, array):
        """Upload on the library's copy-in stream (ibk_grid_upload_async).  `array` must stay alive and
        unmodified until transfers_wait(); page-locked memory is needed for the copy to be asynchronous."""
        a = array
        assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous
        assert a.shape == self.side_shape(patch, axis), (a.shape, self.side_shape(patch, axis))
        self.ctx.check(self.ctx.lib.ibk_grid_upload_async(self.ctx.h, {"u": 0, "f": 1}[which], patch, axis, _dp(a)))

    def grid_download_async(self, which, patch, axis, out):
        """Download into `out` on the copy-out stream (ibk_grid_download_async); valid after transfers_wait()."""
        assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous
        assert out.shape == self.side_shape(patch, axis), (out.shape, self.side_shape(patch, axis))
        self.ctx.check(self.ctx.lib.ibk_grid_download_async(self.ctx.h, {"u": 0, "f": 1}[which], patch, axis, _dp(out)))

    def transfers_wait(self):
        self.ctx.check(self.ctx.lib.ibk_transfers_wait(self.ctx.h))

    def grid_fill(self, which, value):
        self.ctx.check(self.ctx.lib.ibk_grid_fill(self.ctx.h, {"u": 0, "f": 1}[which], float(value)))

    # -- Lagrangian data (LData role) ---------------------------------------------------------------
    def setPositions(self, X):
        X = _f64(X).reshape(-1, self.ndim)
        self.n_markers = X.shape[0]
        self.ctx.check(self.ctx.lib.ibk_markers_set_positions(self.ctx.h, _dp(X), self.n_markers))

    def setLData(self, name, data):
        a = _f64(data).reshape(-1, self.ndim)
        assert a.shape[0] == self.n_markers
        self.ctx.check(self.ctx.lib.ibk_markers_upload(self.ctx.h, COLUMNS[name], _dp(a)))

    def getLData(self, name):
        a = np.zeros((self.n_markers, self.ndim))
        if self.n_markers:
            self.ctx.check(self.ctx.lib.ibk_markers_download(self.ctx.h, COLUMNS[name], _dp(a)))
        return a

    # -- IBStrategy::beginDataRedistribution / endDataRedistribution (IBStrategy.h:455,464) ----------
    def beginDataRedistribution(self):
        self.ctx.check(self.ctx.lib.ibk_rebin(self.ctx.h, int(self.error_if_points_leave_domain)))

    def endDataRedistribution(self):
        pass  # the device rebin does both halves (LDataManager.cpp:1348-1959) in one call

    def getCellsAndOwners(self):
        cells = np.zeros((self.n_markers, self.ndim), dtype=np.int32)
        owner = np.zeros(self.n_markers, dtype=np.int32)
        if self.n_markers:
            self.ctx.check(self.ctx.lib.ibk_bin_get_cells(self.ctx.h, _ip(cells), _ip(owner)))
        return cells, owner

    # -- N1: Lagrangian force + position updates on the device (IBStandardForceGen, IBMethod steps) ----
    def setWallBc(self, acoef, bcoef):
        """Robin coefficients [ndim][2][ndim] of the physical boundaries: switches the fold-back of the spread force on
        (f_phys_bdry_op of IBStrategy::spreadForce; CartSideRobinPhysBdryOp::accumulateFromPhysicalBoundaryData)."""
        a, b = _f64(acoef).reshape(-1), _f64(bcoef).reshape(-1)
        assert a.size == b.size == self.ndim * 2 * self.ndim
        self.ctx.check(self.ctx.lib.ibk_level_set_wall_bc(self.ctx.h, _dp(a), _dp(b)))

    # -- N3: AMR transfer operators between two levels on one device (one IBMethodB200 / context per level) ----
    def prolongFrom(self, coarse, ratio, which="f"):
        """this (finer) level's `which` := CONSERVATIVE_LINEAR_REFINE of the coarser level's: f_prolongation_scheds[ln]->fillData
        before the spread on this level (LDataManager.cpp:611-614).  Returns the number of fine points written."""
        r = np.ascontiguousarray(ratio, dtype=np.int32)
        n = C.c_longlong(0)
        self.ctx.check(self.ctx.lib.ibk_amr_refine_side(self.ctx.h, coarse.ctx.h, 0 if which == "u" else 1, _ip(r), C.byref(n)))
        return n.value

    def coarsenFrom(self, fine, ratio, which="u"):
        """this (coarser) level's `which` := CONSERVATIVE_COARSEN of the finer level's: f_synch_scheds[ln]->coarsenData before
        the interpolation (LDataManager.cpp:728-734).  Returns the number of coarse points written."""
        r = np.ascontiguousarray(ratio, dtype=np.int32)
        n = C.c_longlong(0)
        self.ctx.check(self.ctx.lib.ibk_amr_coarsen_side(self.ctx.h, fine.ctx.h, 0 if which == "u" else 1, _ip(r), C.byref(n)))
        return n.value

    # -- N4: matrix form of the interpolation (PETScMatUtilities::constructPatchLevelSCInterpOp) ----
    def constructInterpOp(self, dof_index, interp_fcn="IB_4"):
        """Rows of the interpolation operator for the resident markers: (cols, vals), each [ndim * n_markers, stencil^ndim],
        row ndim * k + axis.  dof_index[patch][axis]: int32 arrays of side_shape(patch, axis) (SideData<int> with ghosts).
        interp_fcn: "IB_4" / "PIECEWISE_LINEAR" = PETScMatUtilities::ib_4_interp_fcn / pwl_interp_fcn."""
        fcn, S = {"IB_4": (0, 4), "PIECEWISE_LINEAR": (1, 2)}[interp_fcn]
        arrs = [np.ascontiguousarray(dof_index[p][a], dtype=np.int32) for p in range(len(self.boxes)) for a in range(self.ndim)]
        ptrs = (C.POINTER(C.c_int) * len(arrs))(*[_ip(a) for a in arrs])
        rows = self.ndim * self.n_markers
        cols = np.zeros((max(rows, 1), S ** self.ndim), dtype=np.int32)
        vals = np.zeros((max(rows, 1), S ** self.ndim))
        bad = C.c_int(0)
        self.ctx.check(self.ctx.lib.ibk_construct_sc_interp_op(self.ctx.h, fcn, ptrs, _ip(cols), _dp(vals), C.byref(bad)))
        if bad.value:
            raise IBKError(-1, f"constructInterpOp: {bad.value} rows have no local patch that holds their stencil")
        return cols[:rows], vals[:rows]

    def getPatchLists(self, patch):
        """LIndexSetData::cacheLocalIndices for one patch: (lag_idx[n], periodic_shifts[n][ndim], interior[n] as bool)."""
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.ibk_bin_get_patch_lists(self.ctx.h, patch, C.byref(n), None, None, None))
        cnt = n.value
        idx = np.zeros(max(cnt, 1), dtype=np.int32)
        sh = np.zeros((max(cnt, 1), self.ndim))
        interior = np.zeros(max(cnt, 1), dtype=np.int32)
        n = C.c_int(cnt)
        self.ctx.check(self.ctx.lib.ibk_bin_get_patch_lists(self.ctx.h, patch, C.byref(n), _ip(idx), _dp(sh), _ip(interior)))
        return idx[:cnt], sh[:cnt], interior[:cnt].astype(bool)

    def lincomb(self, dst, alpha, a, beta, b):
        """dst = alpha * a + beta * b on marker columns (VecWAXPY / VecAXPBYPCZ of IBMethod.cpp:714-826)."""
        self.ctx.check(self.ctx.lib.ibk_markers_lincomb(self.ctx.h, COLUMNS[dst], float(alpha), COLUMNS[a], float(beta), COLUMNS[b]))

    def setSprings(self, master, slave, kappa, rest_length):
        m, s_ = _i32(master), _i32(slave)
        k, r = _f64(kappa), _f64(rest_length)
        self.ctx.check(self.ctx.lib.ibk_force_set_springs(self.ctx.h, len(m), _ip(m), _ip(s_), _dp(k), _dp(r)))

    def setBeams(self, curr, next_, prev, rigidity, curvature=None):
        c, n, p = _i32(curr), _i32(next_), _i32(prev)
        k = _f64(rigidity)
        cv = _f64(curvature) if curvature is not None else None
        self.ctx.check(self.ctx.lib.ibk_force_set_beams(self.ctx.h, len(c), _ip(c), _ip(n), _ip(p), _dp(k),
                                                        _dp(cv) if cv is not None else None))

    def setTargetPoints(self, idx, kappa, eta, X0):
        i, k, x0 = _i32(idx), _f64(kappa), _f64(X0)
        e = _f64(eta) if eta is not None else None
        self.ctx.check(self.ctx.lib.ibk_force_set_target_points(self.ctx.h, len(i), _ip(i), _dp(k), _dp(e) if e is not None else None,
                                                                _dp(x0)))

    def clearForces(self):
        self.ctx.check(self.ctx.lib.ibk_force_clear(self.ctx.h))

    def computeLagrangianForce(self, x="X", u="U", f="F"):
        """IBMethod::computeLagrangianForce (IBMethod.cpp:834-858): F = springs + beams + target points."""
        self.ctx.check(self.ctx.lib.ibk_compute_lagrangian_force(self.ctx.h, COLUMNS[x], COLUMNS[u], COLUMNS[f]))

    def scaleRows(self, dst, src, ds):
        """dst = src * ds row by row: the F * ds product of LDataManager::spread (LDataManager.cpp:416-447)."""
        w = _f64(ds).reshape(-1)
        assert w.size == self.n_markers
        self.ctx.check(self.ctx.lib.ibk_markers_scale_rows(self.ctx.h, COLUMNS[dst], COLUMNS[src], _dp(w)))

    def resetAnchorPointValues(self, name, anchor_idx):
        """IBMethod::resetAnchorPointValues (IBMethod.cpp:1915-1943): zero the rows of the anchored nodes."""
        a = _i32(anchor_idx)
        self.ctx.check(self.ctx.lib.ibk_markers_zero_rows(self.ctx.h, COLUMNS[name], _ip(a), len(a)))

    def preprocessIntegrateData(self):
        """Start of a step: X_current := the working positions (IBMethod::preprocessIntegrateData keeps X_current)."""
        self.lincomb("X_current", 1.0, "X", 0.0, "X")

    def forwardEulerStep(self, dt):
        """IBMethod::forwardEulerStep (IBMethod.cpp:714-738): X_new = X_current + dt U, then the midpoint data
        X := (X_current + X_new) / 2 (reinitMidpointData, :1900-1912) become the working positions."""
        self.lincomb("X_new", 1.0, "X_current", dt, "U")
        self.lincomb("X", 0.5, "X_current", 0.5, "X_new")

    def midpointStep(self, dt):
        """IBMethod::midpointStep (IBMethod.cpp:768-792): X_new = X_current + dt U_half; midpoint data refreshed."""
        self.lincomb("X_new", 1.0, "X_current", dt, "U")
        self.lincomb("X", 0.5, "X_current", 0.5, "X_new")

    def postprocessIntegrateData(self):
        """End of a step: the new positions become the working and the current ones."""
        self.lincomb("X", 1.0, "X_new", 0.0, "X_new")

    # -- several processes: global Lagrangian indices, marker migration (LDataManager.cpp:1519-1959) --
    def setIds(self, ids, id_bound):
        """Global Lagrangian index of every host row (ibk_markers_set_ids)."""
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        assert a.shape == (self.n_markers,)
        self.ctx.check(self.ctx.lib.ibk_markers_set_ids(self.ctx.h, a.ctypes.data_as(C.POINTER(C.c_uint)), int(id_bound)))

    def getIds(self):
        a = np.zeros(self.n_markers, dtype=np.uint32)
        if self.n_markers:
            self.ctx.check(self.ctx.lib.ibk_markers_get_ids(self.ctx.h, a.ctypes.data_as(C.POINTER(C.c_uint))))
        return a

    def migrate_plan(self, patch_lower, patch_upper, patch_rank, n_ranks, my_rank):
        """Send counts per destination rank for the markers no local patch accepted (ibk_migrate_plan)."""
        lo, hi, rk = _i32(patch_lower).reshape(-1), _i32(patch_upper).reshape(-1), _i32(patch_rank).reshape(-1)
        counts = np.zeros(n_ranks, dtype=np.int32)
        self.ctx.check(self.ctx.lib.ibk_migrate_plan(self.ctx.h, len(rk), _ip(lo), _ip(hi), _ip(rk), int(n_ranks), int(my_rank),
                                                     _ip(counts)))
        return counts

    def migrate_pack(self, device_ptr):
        self.ctx.check(self.ctx.lib.ibk_migrate_pack(self.ctx.h, C.c_void_p(device_ptr)))

    def migrate_unpack(self, device_ptr, n_recv, id_bound):
        self.ctx.check(self.ctx.lib.ibk_migrate_unpack(self.ctx.h, C.c_void_p(device_ptr), int(n_recv), int(id_bound)))
        self.n_markers = int(self.ctx.lib.ibk_markers_count(self.ctx.h))

    def getSortedLagrangianIndices(self):
        lag = np.zeros(self.n_markers, dtype=np.int32)
        if self.n_markers:
            self.ctx.check(self.ctx.lib.ibk_bin_get_order(self.ctx.h, _ip(lag)))
        return lag

    # -- IBStrategy::spreadForce / interpolateVelocity (IBStrategy.h:338-342, 276-280) ----------------
    def spreadForce(self, accumulate_halo=True):
        """f += S[F]: f_data_idx -> the resident f; F/X LData -> the resident columns."""
        self.ctx.check(self.ctx.lib.ibk_spread_force(self.ctx.h, self.spread_kernel_fcn.encode(), int(accumulate_halo)))

    def interpolateVelocity(self, fill_halo=True):
        """U = J[u]: u_data_idx -> the resident u (ghosts filled first when fill_halo)."""
        self.ctx.check(self.ctx.lib.ibk_interpolate_velocity(self.ctx.h, self.interp_kernel_fcn.encode(), int(fill_halo)))

    def spreadForcePart(self, part):
        """ibk_spread_force_part: 1 = interior tiles, 2 = boundary tiles; no halo handling."""
        self.ctx.check(self.ctx.lib.ibk_spread_force_part(self.ctx.h, self.spread_kernel_fcn.encode(), int(part)))

    def interpolateVelocityPart(self, part):
        self.ctx.check(self.ctx.lib.ibk_interpolate_velocity_part(self.ctx.h, self.interp_kernel_fcn.encode(), int(part)))

    def halo(self, which):
        self.ctx.check(self.ctx.lib.ibk_halo_local(self.ctx.h, {"u": 0, "f": 1}[which]))

    def count_touched_dofs(self, kernel_fcn=None):
        t = C.c_longlong()
        self.ctx.check(self.ctx.lib.ibk_count_touched_dofs(self.ctx.h, (kernel_fcn or self.spread_kernel_fcn).encode(),
                                                           C.byref(t)))
        return int(t.value)

    def marker_device_ptr(self, name):
        p, s = C.c_void_p(), C.c_longlong()
        self.ctx.check(self.ctx.lib.ibk_markers_device_ptr(self.ctx.h, {"X": 0, "U": 1, "F": 2}[name], C.byref(p), C.byref(s)))
        return p.value, s.value

    def grid_device_ptr(self, which, patch, axis):
        p, pitch = C.c_void_p(), C.c_longlong()
        dims = (C.c_int * 3)()
        self.ctx.check(self.ctx.lib.ibk_grid_device_ptr(self.ctx.h, {"u": 0, "f": 1}[which], patch, axis, C.byref(p),
                                                        C.byref(pitch), dims))
        return p.value, pitch.value, tuple(dims[d] for d in range(self.ndim))

    def close(self):
        if self.ctx and self.ctx.h:
            self.ctx.lib.ibk_level_destroy(self.ctx.h)


# ------------------------------------------------------------------------------------------------
# N2: IBStandardInitializer's ASCII structure files (src/IB/IBStandardInitializer.cpp), via libibk.so
# ------------------------------------------------------------------------------------------------
class IBStandardInitializer:
    """Reads <base>.vertex / .spring / .beam / .target / .anchor the way IBStandardInitializer::init does
    (IBStandardInitializer.cpp:131-182) for a list of structures on one level; vertex numbers of structure j
    are offset by the vertex counts of the structures before it (:203-210)."""

    def __init__(self, ndim, base_filenames):
        from . import _lib
        lib = _lib.load()
        self.ndim = ndim
        self.X, self.springs, self.beams, self.targets, self.anchors = [], [], [], [], []
        offset = 0

        def check(rc):
            if rc != 0:
                raise IBKError(rc, lib.ibk_io_last_error().decode())

        cnt = C.c_int(0)
        for base in base_filenames:
            path = lambda ext: (base + ext).encode()
            check(lib.ibk_io_read_vertex_file(path(".vertex"), ndim, None, 0, C.byref(cnt)))
            nv = cnt.value
            X = np.zeros((nv, ndim))
            check(lib.ibk_io_read_vertex_file(path(".vertex"), ndim, _dp(X), nv, C.byref(cnt)))
            self.X.append(X)
            # springs
            check(lib.ibk_io_read_spring_file(path(".spring"), nv, offset, None, None, None, None, None, 0, C.byref(cnt)))
            ns = cnt.value
            m, s_, f = np.zeros(ns, np.int32), np.zeros(ns, np.int32), np.zeros(ns, np.int32)
            k, r = np.zeros(ns), np.zeros(ns)
            if ns:
                check(lib.ibk_io_read_spring_file(path(".spring"), nv, offset, _ip(m), _ip(s_), _dp(k), _dp(r), _ip(f), ns, C.byref(cnt)))
            self.springs.append((m, s_, k, r, f))
            # beams
            check(lib.ibk_io_read_beam_file(path(".beam"), nv, offset, ndim, None, None, None, None, None, 0, C.byref(cnt)))
            nb = cnt.value
            pv, cu, nx = np.zeros(nb, np.int32), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
            bend, curv = np.zeros(nb), np.zeros((nb, ndim))
            if nb:
                check(lib.ibk_io_read_beam_file(path(".beam"), nv, offset, ndim, _ip(pv), _ip(cu), _ip(nx), _dp(bend), _dp(curv), nb,
                                                C.byref(cnt)))
            self.beams.append((pv, cu, nx, bend, curv))
            # target points
            check(lib.ibk_io_read_target_file(path(".target"), nv, offset, None, None, None, 0, C.byref(cnt)))
            nt = cnt.value
            ti, tk, te = np.zeros(nt, np.int32), np.zeros(nt), np.zeros(nt)
            if nt:
                check(lib.ibk_io_read_target_file(path(".target"), nv, offset, _ip(ti), _dp(tk), _dp(te), nt, C.byref(cnt)))
            self.targets.append((ti, tk, te))
            # anchors
            check(lib.ibk_io_read_anchor_file(path(".anchor"), nv, offset, None, 0, C.byref(cnt)))
            na = cnt.value
            ai = np.zeros(na, np.int32)
            if na:
                check(lib.ibk_io_read_anchor_file(path(".anchor"), nv, offset, _ip(ai), na, C.byref(cnt)))
            self.anchors.append(ai)
            offset += nv
        self.n_vertices = offset

    def positions(self):
        return np.concatenate(self.X, axis=0)

    def register(self, ib: "IBMethodB200"):
        """Positions, force elements and target positions of all structures into the device-resident method
        (IBMethod::initializeLevelData + IBStandardForceGen::initializeLevelData roles).  Target positions
        are the initial vertex positions (IBTargetPointForceSpec X0 = initial position)."""
        X = self.positions()
        ib.setPositions(X)
        cat = lambda parts, dtype: np.concatenate(parts).astype(dtype) if parts else np.zeros(0, dtype)
        m = cat([s[0] for s in self.springs], np.int32)
        if len(m):
            for s in self.springs:
                if np.any(s[4] != 0):
                    raise IBKError(IBK_ERR_INVALID, "only the default spring force function (index 0) is on the device")
            ib.setSprings(m, cat([s[1] for s in self.springs], np.int32), cat([s[2] for s in self.springs], np.float64),
                          cat([s[3] for s in self.springs], np.float64))
        c = cat([b[1] for b in self.beams], np.int32)
        if len(c):
            ib.setBeams(c, cat([b[2] for b in self.beams], np.int32), cat([b[0] for b in self.beams], np.int32),
                        cat([b[3] for b in self.beams], np.float64), np.concatenate([b[4] for b in self.beams], axis=0))
        t = cat([tt[0] for tt in self.targets], np.int32)
        if len(t):
            ib.setTargetPoints(t, cat([tt[1] for tt in self.targets], np.float64), cat([tt[2] for tt in self.targets], np.float64),
                               X[t])
        return X
