/*
 * ibk.h -- C ABI of the B200-native Lagrangian-Eulerian interaction library (libibk.so).
 *
 * This is the drop-in boundary for IBAMR's IB spread / interpolate hot path.  Every entry
 * point names the reference interface it replaces (file:line under the IBAMR tree); the
 * reference-side bindings a maintainer would add are shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int: IBK_OK (0) or a negative ibk_status; nothing throws across
 *     the boundary; ibk_last_error(ctx) holds the message of the last failure (the reference
 *     aborts through TBOX_ERROR instead, e.g. LEInteractor.cpp:2425-2429, 4491-4498);
 *   - an ibk_ctx belongs to one (process, CUDA device) pair and is not thread-safe, like the
 *     reference (single-threaded MPI ranks);
 *   - all device work is enqueued on the ctx stream (ibk_ctx_set_stream); *_host entry points
 *     copy in/out on that stream and return after the results are in the caller's buffers;
 *   - "h_" pointers are host memory (borrowed for the call), "d_" pointers are device memory;
 *   - grid arrays at the *_host / raw seams use the reference layout: Fortran order,
 *     u(ilower0-g0:iupper0+g0, ilower1-g1:iupper1+g1[, ilower2-g2:iupper2+g2], 0:depth-1)
 *     (SAMRAI ArrayData), marker arrays are AoS [marker][depth] (LData / PETSc block Vec,
 *     ibtk/include/ibtk/LData.h:351-367);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     IBK_ERR_CUDA.
 */
#ifndef IBK_H
#define IBK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define IBK_MAX_DIM 3
#define IBK_MAX_COMP 8

typedef struct ibk_ctx ibk_ctx;

typedef enum ibk_status
{
    IBK_OK = 0,
    IBK_ERR_INVALID = -1,        /* bad argument                                               */
    IBK_ERR_CUDA = -2,           /* CUDA runtime / driver failure (incl. "no device")          */
    IBK_ERR_UNKNOWN_KERNEL = -3, /* LEInteractor.cpp:2100-2104                                  */
    IBK_ERR_GHOST_WIDTH = -4,    /* "insufficient ghost cells", LEInteractor.cpp:4488-4498      */
    IBK_ERR_DEPTH = -5,          /* "side-centered ... requires vector-valued data", :2425-2429 */
    IBK_ERR_STATE = -6,          /* call order violated (e.g. spread before rebin)              */
    IBK_ERR_ESCAPED = -7         /* "IB point has escaped ...", LDataManager.cpp:1410-1416      */
} ibk_status;

/* The delta kernels (names as LEInteractor::string_to_kernel accepts them, LEInteractor.cpp:1923-2016): the five
 * of the hot path's scope and, as a first part of SURVEY.md 8(f) N4, six more that do not depend on the component
 * axis (lagrangian_interaction3d.f.m4: ib_3, bspline_5, bspline_6, piecewise_cubic, ib_5, piecewise_constant). */
typedef enum ibk_kernel
{
    IBK_PIECEWISE_LINEAR = 0,
    IBK_IB_4 = 1,
    IBK_IB_6 = 2,
    IBK_BSPLINE_3 = 3,
    IBK_BSPLINE_4 = 4,
    IBK_IB_3 = 5,
    IBK_BSPLINE_5 = 6,
    IBK_BSPLINE_6 = 7,
    IBK_PIECEWISE_CUBIC = 8,
    IBK_IB_5 = 9,
    IBK_PIECEWISE_CONSTANT = 10,
    /* axis-dependent kernels (the Fortran routines take an `axis` argument, lagrangian_interaction3d.f.m4:237-243,
     * 3510-3516): COMPOSITE_BSPLINE_<A><B> = B-spline of order A along the component's axis, order B elsewhere */
    IBK_COMPOSITE_BSPLINE_32 = 11,
    IBK_COMPOSITE_BSPLINE_23 = 12,
    IBK_COMPOSITE_BSPLINE_43 = 13,
    IBK_COMPOSITE_BSPLINE_34 = 14,
    IBK_COMPOSITE_BSPLINE_54 = 15,
    IBK_COMPOSITE_BSPLINE_45 = 16,
    IBK_COMPOSITE_BSPLINE_65 = 17,
    IBK_COMPOSITE_BSPLINE_56 = 18,
    IBK_DISCONTINUOUS_LINEAR = 19,
    IBK_IB_4_W8 = 20, /* lagrangian_ib_4_w8_*: the 4-point function broadened to 8 meshwidths */
    IBK_KERNEL_LAST = IBK_IB_4_W8,
    IBK_USER_DEFINED = 21 /* "USER_DEFINED": LEInteractor::s_kernel_fcn, see ibk_set_user_kernel */
} ibk_kernel;

/* ---- LEInteractor static queries (ibtk/include/ibtk/LEInteractor.h:99-117) ------------------ */
/* string_to_kernel: returns the ibk_kernel value or IBK_ERR_UNKNOWN_KERNEL. */
int ibk_kernel_from_string(const char* kernel_fcn);
/* LEInteractor::isKnownKernel restricted to the in-scope kernels (LEInteractor.cpp:2038-2050). */
int ibk_is_known_kernel(const char* kernel_fcn);
/* LEInteractor::getStencilSize (LEInteractor.cpp:2052-2108). */
int ibk_get_stencil_size(const char* kernel_fcn);
/* LEInteractor::getMinimumGhostWidth (LEInteractor.cpp:2110-2114). */
int ibk_get_minimum_ghost_width(const char* kernel_fcn);
/* "USER_DEFINED" (LEInteractor.h:82-83: static double (*s_kernel_fcn)(double r), static int s_kernel_fcn_stencil_size; defaults
 * LEInteractor.cpp:2019-2020 = the 4-point function, stencil 4).  Process-wide like the reference's statics; a null function
 * restores the default.  The function is a HOST callback: its values are produced on the host, entry by entry, the way
 * userDefinedInterpolate / userDefinedSpread do (LEInteractor.cpp:6128-6382); the sums over grid data run on the device, the
 * spread in the reference's serial order of additions (no atomics).  Served at the funnel seam (ibk_raw_*) and the patch
 * seams (ibk_{side,cell,node,edge}_*_host); the resident level (ibk_spread_force / ibk_interpolate_velocity) refuses it. */
typedef double (*ibk_kernel_fcn)(double r);
int ibk_set_user_kernel(ibk_kernel_fcn fcn, int stencil_size);

/* ---- context ---------------------------------------------------------------------------------- */
int ibk_ctx_create(int device, ibk_ctx** ctx);
int ibk_ctx_destroy(ibk_ctx* ctx);
const char* ibk_last_error(const ibk_ctx* ctx);
int ibk_ctx_set_stream(ibk_ctx* ctx, void* cuda_stream); /* cudaStream_t */
int ibk_ctx_synchronize(ibk_ctx* ctx);
/* Number of kernels this library launched on the ctx since creation (bench.py's gpu_launches). */
long long ibk_ctx_launch_count(const ibk_ctx* ctx);
/* Device-time (ms) of the last spread / interp / rebin kernel groups, measured with CUDA events
 * on the ctx stream (valid after ibk_ctx_synchronize). which: 0 spread, 1 interp, 2 rebin. */
int ibk_ctx_enable_timing(ibk_ctx* ctx, int enable);
int ibk_ctx_last_ms(ibk_ctx* ctx, int which, float* ms);

/* ---- seam B4: the raw funnel ------------------------------------------------------------------
 * Replaces the Fortran routines lagrangian_<kernel>_{interp,spread}{2,3}d_
 * (ibtk/src/lagrangian/fortran/lagrangian_interaction3d.f.m4:1203-1209, 1344-1350; C++
 * declarations LEInteractor.cpp:237-1518) as called from LEInteractor's private
 * interpolate/spread (LEInteractor.cpp:4470-5230, 5232-6006).  One call handles one array
 * (one SideData axis, or a CellData of `depth` components). */
typedef struct ibk_array_desc
{
    int ndim;                      /* 2 or 3                                                     */
    int depth;                     /* components stored one after another (Fortran last index)   */
    double dx[IBK_MAX_DIM];
    double x_lower[IBK_MAX_DIM];   /* already shifted by -dx/2 on the side axis (LEInteractor.cpp:2464) */
    double x_upper[IBK_MAX_DIM];   /* unused by the in-scope kernels, kept for signature parity  */
    int ilower[IBK_MAX_DIM];       /* data box (toSideBox'ed for SideData)                       */
    int iupper[IBK_MAX_DIM];
    int nugc[IBK_MAX_DIM];         /* ghost width of the array                                   */
    int axis;                      /* `axis` of the axis-dependent routines (LEInteractor passes the SideData /
                                      EdgeData component, 0 otherwise: LEInteractor.h:1341); ignored by the others */
} ibk_array_desc;

/* V(d, indices[l]) = sum_stencil w * u(..., d); markers not listed are left untouched. */
int ibk_raw_interp(ibk_ctx* ctx,
                   int kernel,
                   const ibk_array_desc* desc,
                   const double* d_u,
                   const int* d_indices,
                   const double* d_Xshift, /* [nindices][ndim] or NULL for zeros */
                   int nindices,
                   const double* d_X,      /* AoS [*][ndim] */
                   int n_markers,          /* extent of X / V (max index + 1) */
                   double* d_V);           /* AoS [*][depth] */
/* u(..., d) += w * V(d, indices[l]) / prod(dx), deterministic (no floating-point atomics). */
int ibk_raw_spread(ibk_ctx* ctx,
                   int kernel,
                   const ibk_array_desc* desc,
                   const int* d_indices,
                   const double* d_Xshift,
                   int nindices,
                   const double* d_X,
                   int n_markers,
                   const double* d_V,
                   double* d_u);
/* Same with host buffers (H2D/D2H inside): the signature a Fortran-level shim would bind. */
int ibk_raw_interp_host(ibk_ctx* ctx,
                        int kernel,
                        const ibk_array_desc* desc,
                        const double* h_u,
                        const int* h_indices,
                        const double* h_Xshift,
                        int nindices,
                        const double* h_X,
                        int n_markers,
                        double* h_V);
int ibk_raw_spread_host(ibk_ctx* ctx,
                        int kernel,
                        const ibk_array_desc* desc,
                        const int* h_indices,
                        const double* h_Xshift,
                        int nindices,
                        const double* h_X,
                        int n_markers,
                        const double* h_V,
                        double* h_u);

/* ---- seam B3: patch-level LEInteractor calls --------------------------------------------------
 * What LEInteractor reads from SAMRAI's Patch / CartesianPatchGeometry / SideData. */
typedef struct ibk_patch_desc
{
    int ndim;
    int lower[IBK_MAX_DIM]; /* patch box (cell indices, inclusive)                               */
    int upper[IBK_MAX_DIM];
    int gcw[IBK_MAX_DIM];   /* ghost cell width of the Eulerian data                             */
    double x_lower[IBK_MAX_DIM];
    double x_upper[IBK_MAX_DIM];
    double dx[IBK_MAX_DIM];
    int touches_physical_bdry; /* any getTouchesRegularBoundary() (LEInteractor.cpp:5253-5258)   */
} ibk_patch_desc;

/* Position-only SideData interpolate:
 *   LEInteractor::interpolate(double* Q, int Q_size, int Q_depth, const double* X, int X_size,
 *                             int X_depth, Pointer<SideData> q, Pointer<Patch>, const Box& box,
 *                             const std::string& fcn)     LEInteractor.h:566-575, .cpp:3045-3127
 * h_q[axis] is SideData::getPointer(axis).  Only markers whose cell (getCellIndex) lies in
 * `box` are interpolated; the others keep their Q. */
int ibk_side_interpolate_host(ibk_ctx* ctx,
                              const char* interp_fcn,
                              const ibk_patch_desc* patch,
                              const double* const* h_q,
                              int q_depth,
                              const int* box_lower,
                              const int* box_upper,
                              const double* h_X,
                              int X_size,
                              int X_depth,
                              double* h_Q,
                              int Q_size,
                              int Q_depth);
/* Position-only SideData spread (LEInteractor.h:1132-1141, .cpp:4188-4263): q += S[Q]. */
int ibk_side_spread_host(ibk_ctx* ctx,
                         const char* spread_fcn,
                         const ibk_patch_desc* patch,
                         double* const* h_q,
                         int q_depth,
                         const int* box_lower,
                         const int* box_upper,
                         const double* h_X,
                         int X_size,
                         int X_depth,
                         const double* h_Q,
                         int Q_size,
                         int Q_depth);
/* Position-only CellData forms (LEInteractor.cpp:2805-2863, 3951-4009); q has q_depth comps. */
int ibk_cell_interpolate_host(ibk_ctx* ctx,
                              const char* interp_fcn,
                              const ibk_patch_desc* patch,
                              const double* h_q,
                              int q_depth,
                              const int* box_lower,
                              const int* box_upper,
                              const double* h_X,
                              int X_size,
                              int X_depth,
                              double* h_Q,
                              int Q_size,
                              int Q_depth);
int ibk_cell_spread_host(ibk_ctx* ctx,
                         const char* spread_fcn,
                         const ibk_patch_desc* patch,
                         double* h_q,
                         int q_depth,
                         const int* box_lower,
                         const int* box_upper,
                         const double* h_X,
                         int X_size,
                         int X_depth,
                         const double* h_Q,
                         int Q_size,
                         int Q_depth);
/* NodeData (one array over toNodeBox(box), any depth, index i at x_lower + i dx in every dimension; LEInteractor.cpp:2983-3043,
 * 4122-4186) and EdgeData (one array per axis over toEdgeBox(box, axis), shifted in every dimension but the axis, vector-valued
 * Lagrangian data; LEInteractor.cpp:3260-3340, 4386-4466), position-only forms.  N4 of SURVEY.md 8(f). */
int ibk_node_interpolate_host(ibk_ctx* ctx, const char* interp_fcn, const ibk_patch_desc* patch, const double* h_q, int q_depth,
                              const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth, double* h_Q,
                              int Q_size, int Q_depth);
int ibk_node_spread_host(ibk_ctx* ctx, const char* spread_fcn, const ibk_patch_desc* patch, double* h_q, int q_depth,
                         const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth, const double* h_Q,
                         int Q_size, int Q_depth);
int ibk_edge_interpolate_host(ibk_ctx* ctx, const char* interp_fcn, const ibk_patch_desc* patch, const double* const* h_q,
                              int q_depth, const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth,
                              double* h_Q, int Q_size, int Q_depth);
int ibk_edge_spread_host(ibk_ctx* ctx, const char* spread_fcn, const ibk_patch_desc* patch, double* const* h_q, int q_depth,
                         const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth, const double* h_Q,
                         int Q_size, int Q_depth);
/* Index-set SideData forms (LEInteractor.h:184-192, 704-712; .cpp:2402-2489, 3627-3714): the
 * caller passes the flat lists LIndexSetData caches (local PETSc indices + periodic shifts). */
int ibk_side_interpolate_indexed_host(ibk_ctx* ctx,
                                      const char* interp_fcn,
                                      const ibk_patch_desc* patch,
                                      const double* const* h_q,
                                      const int* h_local_indices,
                                      const double* h_periodic_shifts,
                                      int n_indices,
                                      const double* h_X,
                                      int n_markers,
                                      double* h_Q);
int ibk_side_spread_indexed_host(ibk_ctx* ctx,
                                 const char* spread_fcn,
                                 const ibk_patch_desc* patch,
                                 double* const* h_q,
                                 const int* h_local_indices,
                                 const double* h_periodic_shifts,
                                 int n_indices,
                                 const double* h_X,
                                 int n_markers,
                                 const double* h_Q);

/* ---- seams B1/B2: device-resident level (LDataManager + LData + LIndexSetData roles) --------- */
typedef struct ibk_level_desc
{
    int ndim;
    int n_patches;                      /* patches owned by THIS process on this level           */
    int domain_lower[IBK_MAX_DIM];      /* level index space of the physical domain              */
    int domain_upper[IBK_MAX_DIM];
    double x_lower[IBK_MAX_DIM];        /* physical domain (CartesianGridGeometry)               */
    double x_upper[IBK_MAX_DIM];
    int periodic[IBK_MAX_DIM];          /* periodic_shift != 0                                   */
    int gcw[IBK_MAX_DIM];               /* ghost width of u / f ("ib_ghosts",
                                           src/IB/IBHierarchyIntegrator.cpp:297-303)             */
    const int* patch_lower;             /* [n_patches][ndim]                                     */
    const int* patch_upper;             /* [n_patches][ndim]                                     */
} ibk_level_desc;

/* Registers the level and allocates device-resident side-centred u and f (all patches, with
 * ghosts; device layout is pitched).  Replaces the SAMRAI patch data the
 * integrator allocates for d_u_idx / d_f_idx. */
int ibk_level_create(ibk_ctx* ctx, const ibk_level_desc* desc);
int ibk_level_destroy(ibk_ctx* ctx);

/* Grid data movement (SideData::getPointer(axis) layout on the host side).  which: 0 = u, 1 = f. */
int ibk_grid_upload(ibk_ctx* ctx, int which, int patch, int axis, const double* h_data);
int ibk_grid_download(ibk_ctx* ctx, int which, int patch, int axis, double* h_data);
int ibk_grid_fill(ibk_ctx* ctx, int which, double value); /* HierarchyDataOpsReal::setToScalar */
/* The same transfers on the library's own copy streams (one per direction), so the PCIe traffic of one
 * array overlaps the kernels working on another and the traffic in the opposite direction.  h_data should be
 * page-locked (cudaHostRegister / cudaHostAlloc) or the copy degrades to a synchronous one.  Ordering is kept
 * by the library: an upload starts after the kernels already queued that read the array; every later call
 * that touches the array waits (on the device) for its transfers in flight.  The HOST may read a downloaded
 * array, or reuse an uploaded one, only after ibk_transfers_wait.  These replace the copies SAMRAI's
 * schedules make between the integrator's u/f patch data and the IB scratch data
 * (src/IB/IBHierarchyIntegrator.cpp:300-303 registers d_u_idx / d_f_idx, :366-377 their fill schedules). */
int ibk_grid_upload_async(ibk_ctx* ctx, int which, int patch, int axis, const double* h_data);
int ibk_grid_download_async(ibk_ctx* ctx, int which, int patch, int axis, double* h_data);
int ibk_transfers_wait(ibk_ctx* ctx);

/* LData role: marker columns, AoS on the host side, SoA fp64 on the device.
 * ibk_markers_set_positions replaces LData("X") setup (LDataManager.cpp:2187-2197) and resets
 * the Lagrangian numbering to 0..n-1. */
int ibk_markers_set_positions(ibk_ctx* ctx, const double* h_X, int n_markers);
/* which: 0 = X, 1 = U, 2 = F (and the optional columns 3 = X_current, 4 = X_new, 5 = auxiliary of the N1
 * section below, allocated on first use); AoS [n][ndim] in LAGRANGIAN index order. */
int ibk_markers_upload(ibk_ctx* ctx, int which, const double* h_data);
int ibk_markers_download(ibk_ctx* ctx, int which, double* h_data);
int ibk_markers_count(const ibk_ctx* ctx);

/* LDataManager::beginDataRedistribution + endDataRedistribution (LDataManager.cpp:1348-1959):
 * wrap/clamp X into the domain, cell = getCellIndex(X, grid_geom, ratio), owner patch, stable
 * device radix sort by (patch, brick, cell) with the Lagrangian index as tie-break, permute all
 * marker columns.  error_if_points_leave_domain follows IBMethod's flag (IBMethod.cpp:2060). */
int ibk_rebin(ibk_ctx* ctx, int error_if_points_leave_domain);
/* How many markers a local patch accepted at the last ibk_rebin.  The rest (ibk_markers_count - n_owned: their cell lies in no
 * local patch) are kept behind the binned ones and skipped by every operation until ibk_migrate sends them to their owners;
 * with one process a non-zero rest means the patches do not cover the structure. */
int ibk_markers_owned_count(ibk_ctx* ctx, int* n_owned);

/* ---- more than one process: global Lagrangian indices and marker migration ------------------------------
 * With one process the host rows of ibk_markers_upload/download ARE the Lagrangian indices 0..n-1.  With
 * several, each process holds a subset: ibk_markers_set_ids attaches the global Lagrangian index of every
 * host row (h_ids[n], all < id_bound); the in-cell order of the binning then follows the global index
 * (LDataManager.cpp:1505).  ibk_markers_get_ids returns them in host-row order.
 *
 * Migration = LDataManager::endDataRedistribution's scatter of every LData row to the process that owns the
 * marker's new cell (LDataManager.cpp:1519-1959, VecScatter :1824-1837).  After ibk_rebin the markers that no
 * local patch accepts sit behind the binned ones; then
 *   ibk_migrate_plan    finds their destination rank from the level's global box list (patch_lower/upper:
 *                       [n_patches][ndim] cell boxes, patch_rank[n_patches]) and fills h_send_counts[n_ranks];
 *                       IBK_ERR_ESCAPED if a marker lies in no patch of the level;
 *   ibk_migrate_pack    writes sum(h_send_counts) rows of 3*ndim+1 doubles [X, U, F, index] into the DEVICE
 *                       buffer d_buf, grouped by destination rank in ascending order (the send buffer of an
 *                       all-to-all: MPI_Alltoallv / ncclSend+ncclRecv; ibamr_b200/halo.py::MarkerMigration);
 *   ibk_migrate_unpack  drops the markers that left, appends the n_recv rows that arrived, and renumbers the
 *                       host rows: ascending global index among the markers now held.  ibk_rebin must follow. */
int ibk_markers_set_ids(ibk_ctx* ctx, const unsigned* h_ids, unsigned id_bound);
int ibk_markers_get_ids(ibk_ctx* ctx, unsigned* h_ids);
int ibk_migrate_plan(ibk_ctx* ctx, int n_patches, const int* patch_lower, const int* patch_upper, const int* patch_rank,
                     int n_ranks, int my_rank, int* h_send_counts);
int ibk_migrate_pack(ibk_ctx* ctx, double* d_buf);
int ibk_migrate_unpack(ibk_ctx* ctx, const double* d_buf, int n_recv, unsigned id_bound);

/* ---- N1: Lagrangian forces and position updates with the markers resident -------------
 * Marker columns: IBK_COL_X is the one spread / interpolate / re-bin work on (IBMethod's X_LE / half-time data).
 *
 * ibk_markers_lincomb   dst = alpha * a + beta * b on whole columns: the VecWAXPY / VecAXPBYPCZ calls of
 *                       IBMethod::forwardEulerStep / midpointStep / trapezoidalStep / reinitMidpointData
 *                       (src/IB/IBMethod.cpp:714-826, 1900-1912).  Writing IBK_COL_X un-bins the markers.
 * ibk_markers_zero_rows zeroes the rows of the listed Lagrangian indices: IBMethod::resetAnchorPointValues
 *                       (IBMethod.cpp:1915-1943).
 * ibk_force_set_*       the force elements IBStandardForceGen gathers from the node specs
 *                       (IBStandardForceGen.cpp:715-811 springs, 933-1035 beams, 1150-1199 target points), by
 *                       Lagrangian index; springs use the default force function kappa * (R - rest_length)
 *                       (IBSpringForceFunctions.h:99-103); eta / curvature may be NULL (zero).  Set them after
 *                       the markers (and after ibk_markers_set_ids, if used); ibk_force_clear drops them all.
 * ibk_compute_lagrangian_force  F(f_col) = springs + beams + target points at positions x_col, velocities u_col:
 *                       IBMethod::computeLagrangianForce (IBMethod.cpp:834-858) over
 *                       IBStandardForceGen::computeLagrangianForce (IBStandardForceGen.cpp:253-303).  Every node
 *                       gathers its elements in the reference's order: no atomics, bit-reproducible.  With
 *                       several processes every element's nodes must be on one rank (IBK_ERR_STATE otherwise). */
enum
{
    IBK_COL_X = 0,
    IBK_COL_U = 1,
    IBK_COL_F = 2,
    IBK_COL_X_CURRENT = 3,
    IBK_COL_X_NEW = 4,
    IBK_COL_AUX = 5
};
int ibk_markers_lincomb(ibk_ctx* ctx, int dst, double alpha, int a, double beta, int b);
int ibk_markers_zero_rows(ibk_ctx* ctx, int which, const int* lag_idx, int n);
/* dst = src * ds row by row (h_ds[n], host-row order): the F * ds product LDataManager::spread forms when it is given a
 * node-weight LData (LDataManager.cpp:416-447). */
int ibk_markers_scale_rows(ibk_ctx* ctx, int dst, int src, const double* h_ds);
int ibk_force_set_springs(ibk_ctx* ctx, int n, const int* master, const int* slave, const double* kappa, const double* rest_length);
int ibk_force_set_beams(ibk_ctx* ctx, int n, const int* curr, const int* next, const int* prev, const double* rigidity,
                        const double* curvature);
int ibk_force_set_target_points(ibk_ctx* ctx, int n, const int* idx, const double* kappa, const double* eta, const double* X0);
int ibk_force_clear(ibk_ctx* ctx);
int ibk_compute_lagrangian_force(ibk_ctx* ctx, int x_col, int u_col, int f_col);

/* ---- N2: the ASCII structure files of IBStandardInitializer (host only, no context) -----
 * Grammar and validity rules of src/IB/IBStandardInitializer.cpp (readVertexFiles :184-294, readSpringFiles
 * :297-528, readBeamFiles :766-1002, readTargetPointFiles :1322-1517, readAnchorPointFiles :1520-1643): comments
 * after '!', '#', '%'; line 1 = number of entries; indices in [0, n_vertices) and returned with vertex_offset
 * added; duplicates skipped; springs stored smaller index first.  Output arrays may be NULL (count only) and hold
 * at most `capacity` entries; *n_* receives the number of entries kept.  A missing .vertex file is an error, the
 * other files are optional (count 0).  Errors: IBK_ERR_INVALID with the text in ibk_io_last_error(). */
const char* ibk_io_last_error(void);
int ibk_io_read_vertex_file(const char* path, int ndim, double* X, int capacity, int* n_vertices);
int ibk_io_read_spring_file(const char* path, int n_vertices, int vertex_offset, int* master, int* slave, double* kappa,
                            double* rest_length, int* force_fcn_idx, int capacity, int* n_springs);
int ibk_io_read_beam_file(const char* path, int n_vertices, int vertex_offset, int ndim, int* prev, int* curr, int* next,
                          double* rigidity, double* curvature, int capacity, int* n_beams);
int ibk_io_read_target_file(const char* path, int n_vertices, int vertex_offset, int* idx, double* kappa, double* eta, int capacity,
                            int* n_targets);
int ibk_io_read_anchor_file(const char* path, int n_vertices, int vertex_offset, int* idx, int capacity, int* n_anchors);


/* Binning products for parity checks (LIndexSetData role), in LAGRANGIAN index order:
 * cells [n][ndim] (level cell index), owner [n] (local patch number or -1). */
int ibk_bin_get_cells(ibk_ctx* ctx, int* h_cells, int* h_owner);
/* Sorted order: h_lag_idx[i] = Lagrangian index of the marker stored at sorted position i
 * (the "local PETSc index -> Lagrangian index" map, LDataManager.cpp:2897-2911). */
int ibk_bin_get_order(ibk_ctx* ctx, int* h_lag_idx);
/* The index sets of LIndexSetData::cacheLocalIndices for one patch (ibtk/src/lagrangian/LIndexSetData.cpp:53-141): every
 * marker whose cell, or a periodic image of it, lies in the patch's ghost box, in the reference's order (cell k-j-i inside the
 * ghost box, then Lagrangian index), with the periodic shift of the image (:89-101, [n][ndim]) and interior (1) / ghost (0)
 * (:104).  The interior / ghost lists of the reference are the sublists by that flag.  *n_entries: in = capacity of the
 * arrays, out = length of the list; null arrays or too small a capacity: the count only.  The device holds the markers this
 * process owns: the lists are complete when all patches of the level are local. */
int ibk_bin_get_patch_lists(ibk_ctx* ctx, int patch, int* n_entries, int* h_lag_idx, double* h_shift, int* h_interior);

/* LDataManager::spread core (LDataManager.cpp:551-667) as IBMethod::spreadForce calls it
 * (src/IB/IBMethod.cpp:972-995): f += S[F] from the markers each patch OWNS into interior and
 * ghost cells, followed (accumulate_halo != 0) by the ghost-region sum onto the owning DOFs
 * among this process's patches incl. periodic wrap (SAMRAIGhostDataAccumulator semantics,
 * ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:295-353). */
int ibk_spread_force(ibk_ctx* ctx, const char* spread_fcn, int accumulate_halo);
/* The pieces of accumulate_halo != 0, for callers that interleave an inter-process exchange
 * (ibamr_b200/halo.py): ibk_spread_begin zeroes the ghost regions of f and parks + zeroes the
 * pre-existing content of the shared boundary face layers (the reference spreads into a zeroed f and
 * adds the old f afterwards, LDataManager.cpp:589-594, 662-663); then ibk_spread_force(.., 0),
 * pack for remote ranks, ibk_halo_local(ctx, 1), remote unpack-adds; ibk_spread_end restores the
 * parked face content. */
int ibk_spread_begin(ibk_ctx* ctx);
int ibk_spread_end(ibk_ctx* ctx);
/* Physical boundaries (N3 of SURVEY.md 8(f), the f_phys_bdry_op argument of IBStrategy::spreadForce): what the spread puts
 * into ghost cells OUTSIDE a non-periodic domain is folded back into the interior by the adjoint of the linear ghost-cell
 * extrapolation of the Robin condition a u + b du/dn = g, CartSideRobinPhysBdryOp::accumulateFromPhysicalBoundaryData
 * (ibtk/src/boundary/physical_boundary/CartSideRobinPhysBdryOp.cpp:552-617, called LDataManager.cpp:653-657; kernels
 * fortran/cartphysbdryop3d.f.m4:78-168 transverse components, :787-905 normal component, adjoint_op = 1, homogeneous).
 * acoef / bcoef: [ndim][2 sides][ndim components].  Built for walls in ONE dimension (the others periodic): with walls in
 * two or three dimensions the co-dimension two / three extrapolations would be needed and the call is refused.
 * Null pointers switch the fold-back off (the ghost values outside the domain are then dropped).
 * ibk_spread_fold_walls runs it (once per spread; ibk_spread_force, ibk_halo_local(f) and ibk_halo_accumulate_post call it). */
int ibk_level_set_wall_bc(ibk_ctx* ctx, const double* acoef, const double* bcoef);
int ibk_spread_fold_walls(ibk_ctx* ctx);
/* AMR transfer operators either side of the path (N3 of SURVEY.md 8(f)), side-centred data, between two levels registered
 * on the SAME device (one context per level; `fine`'s index space is `coarse`'s refined by ratio[ndim]).
 *   ibk_amr_refine_side(fine, coarse, 1, ratio):  f_prolongation_scheds[ln]->fillData before the spread on level ln
 *     (LDataManager.cpp:611-614; "CONSERVATIVE_LINEAR_REFINE", src/IB/IBHierarchyIntegrator.cpp:374-377): every point of the
 *     fine arrays, ghosts included, whose coarse stencil lies inside a coarse patch's array (fill the coarse ghosts first);
 *   ibk_amr_coarsen_side(coarse, fine, 0, ratio): f_synch_scheds[ln]->coarsenData before the interpolation, finest level
 *     first (LDataManager.cpp:728-734; "CONSERVATIVE_COARSEN", IBHierarchyIntegrator.cpp:369-372): the coarse patches' own
 *     sides that the own sides of a fine patch tile.
 * The operators are SAMRAI's CartesianSideDoubleConservativeLinearRefine / CartesianSideDoubleWeightedAverage (third party,
 * not in the reference tree): restated from their published algorithm, see csrc/ibk_amr.cu.  which: 0 = u, 1 = f.
 * n_points (may be null): points written.  The work runs on the destination context's stream, ordered after everything
 * queued on the source context's stream. */
/* Matrix form of the interpolation (N4 of SURVEY.md 8(f)): PETScMatUtilities::constructPatchLevelSCInterpOp
 * (ibtk/src/math/PETScMatUtilities.cpp:783-1020), the operator IBMethod::constructInterpOp gives the implicit solver
 * (src/IB/IBMethod.cpp:1030-1050).  The reference fills a PETSc AIJ matrix (third party) by one MatSetValues call per row;
 * here the rows of the resident markers are computed on the device and returned as arrays with a fixed row length
 * stencil^ndim (CSR with row_ptr[r] = r * stencil^ndim): h_cols / h_vals [ndim * n_markers][stencil^ndim], row ndim * k + axis
 * for the marker of host row k, entries in the reference's box-iterator order (x fastest).
 * interp_fcn: PETScMatUtilities::ib_4_interp_fcn (stencil 4) or pwl_interp_fcn (stencil 2) (PETScMatUtilities.h:156-176).
 * h_dof_index[patch * ndim + axis]: the SideData<int> DOF numbers of (patch, axis), dense, x fastest, ghost width = the level's.
 * *n_unplaced (may be null): rows whose marker lies in no local patch or whose stencil leaves its ghost box (the reference
 * asserts there are none); their columns are -1. */
enum
{
    IBK_INTERP_FCN_IB_4 = 0,
    IBK_INTERP_FCN_PWL = 1
};
int ibk_construct_sc_interp_op(ibk_ctx* ctx, int interp_fcn, const int* const* h_dof_index, int* h_cols, double* h_vals,
                               int* n_unplaced);
int ibk_amr_refine_side(ibk_ctx* fine, ibk_ctx* coarse, int which, const int* ratio, long long* n_points);
int ibk_amr_coarsen_side(ibk_ctx* coarse, ibk_ctx* fine, int which, const int* ratio, long long* n_points);
/* LDataManager::interp core (LDataManager.cpp:698-813) as IBMethod::interpolateVelocity calls it
 * (IBMethod.cpp:672-694): (fill_halo != 0) ghost fill of u among this process's patches incl.
 * periodic wrap (replaces u_ghost_fill_scheds[ln]->fillData, :744), then U = J[u]. */
int ibk_interpolate_velocity(ibk_ctx* ctx, const char* interp_fcn, int fill_halo);
/* The same operations restricted to a part of the marker tiles, without any halo handling: part 1 = the tiles that
 * touch neither ghost cells nor the layers next to a patch boundary (nothing an exchange reads or writes), part 2 =
 * the others, part 0 = all.  They let the inter-process exchange overlap the bulk of the work
 * (ibamr_b200/halo.py::HaloExchange.*_post / *_finish, bench.py):
 *   spread:  ibk_spread_begin; part 2; ibk_halo_pack ...; post the exchange; part 1; ibk_halo_local(f);
 *            complete the exchange; ibk_halo_unpack(add) ...; ibk_spread_end
 *   interp:  ibk_halo_pack ...; post the exchange; ibk_halo_local(u); part 1; complete the exchange;
 *            ibk_halo_unpack(copy) ...; part 2 */
int ibk_spread_force_part(ibk_ctx* ctx, const char* spread_fcn, int part);
int ibk_interpolate_velocity_part(ibk_ctx* ctx, const char* interp_fcn, int part);

/* Halo ops on their own (device pack/unpack around an external exchange; multi-process runs
 * move the packed buffers with NCCL, see ibamr_b200/halo.py).  which: 0 = u (copy), 1 = f (add). */
int ibk_halo_local(ibk_ctx* ctx, int which);
/* Region pack/unpack for inter-process exchange.  Regions are given in the array index space
 * of (patch, axis): lower/upper inclusive.  mode for unpack: 0 = copy, 1 = add. */
int ibk_halo_pack(ibk_ctx* ctx, int which, int patch, int axis, const int* lower, const int* upper, double* d_buf);
int ibk_halo_unpack(ibk_ctx* ctx, int which, int patch, int axis, const int* lower, const int* upper,
                    const double* d_buf, int mode);
/* All regions of one neighbour's message in one call (items in buffer order; buf_offset[k] in doubles). */
int ibk_halo_pack_many(ibk_ctx* ctx, int which, int n_items, const int* patch, const int* axis, const int* lower, const int* upper,
                       const long long* buf_offset, double* d_buf);
int ibk_halo_unpack_many(ibk_ctx* ctx, int which, int n_items, const int* patch, const int* axis, const int* lower,
                         const int* upper, const long long* buf_offset, const double* d_buf, int mode);

/* ---- the multi-rank layer (one process per GPU; seam B1 across processes) -------------------------------------
 * What the reference reaches from C++ around the hot path when the level is spread over MPI ranks:
 *   fill        ghost cells of u <- owner interiors, periodic wrap included
 *               (u_ghost_fill_scheds[ln]->fillData, ibtk/src/lagrangian/LDataManager.cpp:744);
 *   accumulate  owner interiors of f += every other copy of the DOF, ghost copies and the interior copy of a face shared
 *               by two patches (SAMRAIGhostDataAccumulator::accumulateGhostData,
 *               ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:327-344, called LDataManager.cpp:597-620);
 *   migrate     the scatter of the marker rows to their new owners (LDataManager.cpp:1824-1837).
 *
 * PLAN (host code, no context, no GPU): every rank derives the same plan from the global patch list, so no metadata is
 * exchanged.  table 0 = fill, 1 = accumulate.  A message is everything one rank sends another in one exchange; its items
 * are regions in a canonical order ((axis, destination patch, source rank, source patch, region)), unpack-adds run in
 * ascending source rank, then item order: sums are reproducible.  Boxes are [n][ndim] cell boxes, inclusive;
 * src_local / dst_local number a patch among the patches of its own rank, in list order. */
typedef struct ibk_halo_plan ibk_halo_plan;
int ibk_halo_plan_create(int ndim, int n_patches, const int* lower, const int* upper, const int* rank, const int* domain_ncells,
                         const int* periodic, const int* gcw, int my_rank, ibk_halo_plan** out);
void ibk_halo_plan_destroy(ibk_halo_plan* plan);
int ibk_halo_plan_messages(const ibk_halo_plan* plan, int table); /* number of messages that involve my_rank (< 0: error) */
int ibk_halo_plan_message(const ibk_halo_plan* plan, int table, int k, int* src_rank, int* dst_rank, int* n_items, long long* count);
int ibk_halo_plan_items(const ibk_halo_plan* plan, int table, int k, int* axis, int* src_local, int* dst_local, int* src_lo, int* src_hi,
                        int* dst_lo, int* dst_hi);
/* COMMUNICATOR of a context.  NCCL (resolved with dlopen at the first call: libibk.so itself does not link it): rank 0 obtains
 * a 128-byte id with ibk_comm_unique_id and hands it to the other ranks by whatever the host program has (MPI_Bcast in
 * IBAMR), every rank calls ibk_comm_init.  ibk_comm_init_loopback makes several contexts of ONE process the ranks
 * 0..nranks-1 of a communicator that moves messages by device copies (tests on one GPU; one process driving several GPUs).
 * ibk_comm_set_patches (after ibk_level_create; the same list on every rank, the k-th patch of a rank in the list being
 * the k-th patch of its level) builds the plan and allocates the message buffers. */
int ibk_comm_unique_id(void* id128);
int ibk_comm_init(ibk_ctx* ctx, const void* id128, int rank, int nranks);
int ibk_comm_init_loopback(ibk_ctx** ctxs, int nranks);
int ibk_comm_set_patches(ibk_ctx* ctx, int n_patches, const int* lower, const int* upper, const int* rank);
int ibk_comm_destroy(ibk_ctx* ctx);
/* SMs the persistent spread kernel leaves free so that message kernels can start while it runs (ibk_comm_init sets 8 for the
 * NCCL transport, IBK_COMM_RESERVE_SMS overrides; 0 when no message has to start during a spread). */
int ibk_comm_set_reserved_sms(ibk_ctx* ctx, int n_sms);
/* The exchanges, split so that the messages are in flight while the tiles that do not touch the exchanged regions are
 * processed (ibk_*_part): post = pack on the context's stream + start the messages on its communication stream,
 * finish = the context's stream waits for them + unpack (fill: copy, accumulate: add).  Sequences as listed at
 * ibk_spread_force_part.  Every rank posts before any rank can finish (collective, like the schedules they replace). */
int ibk_halo_fill_post(ibk_ctx* ctx);
int ibk_halo_fill_finish(ibk_ctx* ctx);
int ibk_halo_accumulate_post(ibk_ctx* ctx);
int ibk_halo_accumulate_finish(ibk_ctx* ctx);
long long ibk_halo_bytes(ibk_ctx* ctx, int which); /* bytes this rank sends per fill (0) / accumulate (1) */
/* Marker migration over the communicator: ibk_migrate_plan, counts all-gathered, rows sent and received in one group,
 * ibk_migrate_unpack.  ibk_rebin precedes and follows.  ibk_migrate_loopback does it for all ranks of a loopback
 * communicator in one call. */
int ibk_migrate(ibk_ctx* ctx, unsigned id_bound, int* n_sent, int* n_received);
int ibk_migrate_loopback(ibk_ctx** ctxs, int nranks, unsigned id_bound, int* n_moved);

/* Device pointers for zero-copy callers (torch tensors, NCCL): SoA marker columns in SORTED
 * order ([ndim][capacity] with the given stride) and pitched grid arrays. */
int ibk_markers_device_ptr(ibk_ctx* ctx, int which, double** d_ptr, long long* stride);
int ibk_grid_device_ptr(ibk_ctx* ctx, int which, int patch, int axis, double** d_ptr, long long* pitch,
                        int* dims /* [ndim] incl. ghosts */);

/* Algorithmic-byte accounting of SURVEY.md 8(d): number of distinct side DOFs (all axes) in the
 * stencil support of >= 1 marker, computed on the device from the sorted keys. */
int ibk_count_touched_dofs(ibk_ctx* ctx, const char* kernel_fcn, long long* touched);

#ifdef __cplusplus
}
#endif
#endif /* IBK_H */
