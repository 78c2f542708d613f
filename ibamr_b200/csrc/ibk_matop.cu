// ibk_matop.cu -- the matrix form of the interpolation (row N4 of SURVEY 8(f)):
// PETScMatUtilities::constructPatchLevelSCInterpOp (ibtk/src/math/PETScMatUtilities.cpp:783-1020), the operator
// IBMethod::constructInterpOp hands to the implicit solver (src/IB/IBMethod.cpp:1030-1050).
//
// The reference fills a PETSc AIJ matrix (third party) with one MatSetValues call per (marker, axis) row: stencil^NDIM
// columns taken from the patch's SideData<int> of DOF numbers, values = tensor products of the 1-D weights.  Here the rows
// are produced on the device and handed back as arrays (fixed row length, so CSR with row_ptr[r] = r * stencil^NDIM):
// what the maintainer's binding passes to MatSetValues / MatCreateMPIAIJWithArrays.  The two weight functions are the ones
// PETScMatUtilities itself provides and IBAMR uses (PETScMatUtilities.h:156-176): ib_4_interp_fcn, pwl_interp_fcn.
//
//   cell      = IndexUtilities::getCellIndex(X, grid_geom, ratio)                               (:845)
//   X_cell    = (cell - domain_lower + 1/2) dx + x_lower                                        (:850-857)
//   patch     = a local patch that holds the cell, else one that holds it in its first ghost layer (:861-877)
//   lower_d   = cell_d - s/2 + 1 along the row's axis; elsewhere cell_d - s/2 if X_d <= X_cell_d, else cell_d - s/2 + 1 (:901-917)
//   r_d       = (X_d - ((lower_d - domain_lower_d + (d == axis ? 0 : 1/2)) dx_d + x_lower_d)) / dx_d;  w_d = interp_fcn(r_d)  (:978-983)
//   entries   in box-iterator order (x fastest): value = ((1 * w_0) * w_1) * w_2, column = dof_index(i, axis)  (:990-999)
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "ibk_ctx.h"
#include "ibk_device.cuh"
#include "../../include/ibk.h"

namespace ibk
{
int fail(ibk_ctx* ctx, int code, const std::string& msg);
int cuda_fail(ibk_ctx* ctx, cudaError_t e, const char* what);

constexpr int MATOP_MAX_PATCHES = 32; // (the geometry travels as a kernel parameter)
struct MatOpPatch
{
    int lower[3], upper[3];
    const int* dof[3]; // per axis: SideData<int> with the level's ghost width, dense, x fastest
    int n[3][3];       // extents of those arrays
};
struct MatOpGeom
{
    int ndim, stencil, fcn, n_patches;
    double x_lower[3], x_upper[3], dx[3];
    int dom_lo[3], dom_hi[3], gcw[3];
    MatOpPatch patch[MATOP_MAX_PATCHES];
};

template <int S>
__device__ __forceinline__ void matop_weights(int fcn, double r, double* w)
{
    if (S == 4)
    {
        // ib_4_interp_fcn (PETScMatUtilities.h:156-164), the operations in the order C++ evaluates them
        (void)fcn;
        const double q = sqrt(__dsub_rn(__dadd_rn(-7.0, __dmul_rn(12.0, r)), __dmul_rn(__dmul_rn(4.0, r), r)));
        const double a = __dsub_rn(5.0, __dmul_rn(2.0, r)), b = __dadd_rn(-1.0, __dmul_rn(2.0, r));
        w[0] = __dmul_rn(0.125, __dsub_rn(a, q));
        w[1] = __dmul_rn(0.125, __dadd_rn(a, q));
        w[2] = __dmul_rn(0.125, __dadd_rn(b, q));
        w[3] = __dmul_rn(0.125, __dsub_rn(b, q));
    }
    else
    {
        // pwl_interp_fcn (:171-176)
        w[0] = __dsub_rn(1.0, r);
        w[1] = r;
    }
}

template <int S>
__global__ void matop_rows_kernel(const __grid_constant__ MatOpGeom g, const double* __restrict__ X, long long stride,
                                  const uint32_t* __restrict__ lag, int n, int* __restrict__ cols, double* __restrict__ vals,
                                  int* __restrict__ n_unplaced)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * g.ndim) return;
    const int i = t / g.ndim, axis = t % g.ndim;
    const int ndim = g.ndim;
    double x[3] = { 0, 0, 0 }, x_cell[3];
    int cell[3] = { 0, 0, 0 };
    for (int d = 0; d < ndim; ++d)
    {
        x[d] = X[d * stride + i];
        cell[d] = cell_index_1d(x[d], g.x_lower[d], g.x_upper[d], g.dx[d], g.dom_lo[d], g.dom_hi[d]);
        x_cell[d] = __dadd_rn(__dmul_rn(__dadd_rn((double)(cell[d] - g.dom_lo[d]), 0.5), g.dx[d]), g.x_lower[d]);
    }
    int p = -1;
    for (int growth = 0; growth <= 1 && p < 0; ++growth)
        for (int q = 0; q < g.n_patches && p < 0; ++q)
        {
            bool in = true;
            for (int d = 0; d < ndim; ++d) in = in && cell[d] >= g.patch[q].lower[d] - growth && cell[d] <= g.patch[q].upper[d] + growth;
            if (in) p = q;
        }
    int npts = 1;
    for (int d = 0; d < ndim; ++d) npts *= S;
    const long long row = (long long)ndim * lag[i] + axis;
    int* crow = cols + row * npts;
    double* vrow = vals + row * npts;
    if (p < 0)
    {
        atomicAdd(n_unplaced, 1);
        for (int e = 0; e < npts; ++e)
        {
            crow[e] = -1;
            vrow[e] = 0.0;
        }
        return;
    }
    const MatOpPatch& P = g.patch[p];
    int lower[3] = { 0, 0, 0 };
    double w[3][S];
    for (int d = 0; d < ndim; ++d)
    {
        if (d == axis) lower[d] = cell[d] - S / 2 + 1;
        else lower[d] = x[d] <= x_cell[d] ? cell[d] - S / 2 : cell[d] - S / 2 + 1;
        const double x_sl = __dadd_rn(__dmul_rn(__dadd_rn((double)(lower[d] - g.dom_lo[d]), d == axis ? 0.0 : 0.5), g.dx[d]), g.x_lower[d]);
        matop_weights<S>(g.fcn, __ddiv_rn(__dsub_rn(x[d], x_sl), g.dx[d]), w[d]);
    }
    // the stencil box must lie in the ghost box of the patch's side data (the reference asserts it, :929)
    bool inside = true;
    for (int d = 0; d < ndim; ++d)
        inside = inside && lower[d] >= P.lower[d] - g.gcw[d] && lower[d] + S - 1 <= P.upper[d] + g.gcw[d] + (d == axis ? 1 : 0);
    if (!inside)
    {
        atomicAdd(n_unplaced, 1);
        for (int e = 0; e < npts; ++e)
        {
            crow[e] = -1;
            vrow[e] = 0.0;
        }
        return;
    }
    const int* dof = P.dof[axis];
    const int n0 = P.n[axis][0], n1 = P.n[axis][1];
    int e = 0;
    for (int kz = 0; kz < (ndim == 3 ? S : 1); ++kz)
        for (int ky = 0; ky < S; ++ky)
            for (int kx = 0; kx < S; ++kx, ++e)
            {
                double v = __dmul_rn(w[0][kx], w[1][ky]);
                if (ndim == 3) v = __dmul_rn(v, w[2][kz]);
                const int j0 = lower[0] + kx - (P.lower[0] - g.gcw[0]);
                const int j1 = lower[1] + ky - (P.lower[1] - g.gcw[1]);
                const int j2 = ndim == 3 ? lower[2] + kz - (P.lower[2] - g.gcw[2]) : 0;
                crow[e] = dof[((long long)j2 * n1 + j1) * n0 + j0];
                vrow[e] = v;
            }
}
} // namespace ibk

using namespace ibk;

#define CK(call)                                                   \
    do                                                             \
    {                                                              \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// h_dof_index[patch * ndim + axis]: the SideData<int> DOF numbers of (patch, axis), dense, x fastest, with the level's ghost
// width (the array u has).  h_cols / h_vals: [ndim * n_markers][stencil^ndim], row ndim * k + axis for the marker of host
// row k.  Rows whose marker lies in no local patch (or whose stencil leaves the patch's ghost box) get column -1 and are
// counted in *n_unplaced (the reference asserts that there are none).
extern "C" int ibk_construct_sc_interp_op(ibk_ctx* ctx, int interp_fcn, const int* const* h_dof_index, int* h_cols, double* h_vals,
                                          int* n_unplaced)
{
    if (!ctx) return IBK_ERR_INVALID;
    LevelState& lv = ctx->lv;
    if (!lv.valid) return fail(ctx, IBK_ERR_STATE, "no level registered (ibk_level_create)");
    if (interp_fcn != IBK_INTERP_FCN_IB_4 && interp_fcn != IBK_INTERP_FCN_PWL)
        return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, "interp_fcn must be IBK_INTERP_FCN_IB_4 or IBK_INTERP_FCN_PWL");
    if (!h_dof_index || !h_cols || !h_vals) return fail(ctx, IBK_ERR_INVALID, "null pointer");
    if ((int)lv.patches.size() > MATOP_MAX_PATCHES) return fail(ctx, IBK_ERR_INVALID, "too many patches for the matrix-form operator");
    const int ndim = lv.ndim, S = interp_fcn == IBK_INTERP_FCN_IB_4 ? 4 : 2;
    for (int d = 0; d < ndim; ++d)
        if (lv.gcw[d] < S / 2) return fail(ctx, IBK_ERR_GHOST_WIDTH, "the ghost width of the level is smaller than half the stencil");
    if (n_unplaced) *n_unplaced = 0;
    const int n = lv.n;
    if (n == 0) return IBK_OK;
    CK(cudaSetDevice(ctx->device));
    MatOpGeom g;
    std::memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    g.stencil = S;
    g.fcn = interp_fcn;
    g.n_patches = (int)lv.patches.size();
    for (int d = 0; d < 3; ++d)
    {
        g.x_lower[d] = lv.x_lower[d];
        g.x_upper[d] = lv.x_upper[d];
        g.dx[d] = lv.dx[d];
        g.dom_lo[d] = lv.domain_lower[d];
        g.dom_hi[d] = lv.domain_upper[d];
        g.gcw[d] = lv.gcw[d];
    }
    // DOF numbers: one device buffer, arrays back to back
    size_t total = 0;
    for (const PatchState& ps : lv.patches)
        for (int a = 0; a < ndim; ++a) total += (size_t)ps.n[a][0] * ps.n[a][1] * ps.n[a][2];
    int npts = 1;
    for (int d = 0; d < ndim; ++d) npts *= S;
    const size_t rows = (size_t)ndim * n;
    CK(ctx->b_user[0].reserve(sizeof(int) * total));
    CK(ctx->b_user[1].reserve(sizeof(int) * (rows * npts + 1)));
    CK(ctx->b_user[2].reserve(sizeof(double) * rows * npts));
    int* d_dof = ctx->b_user[0].as<int>();
    size_t off = 0;
    for (size_t p = 0; p < lv.patches.size(); ++p)
    {
        const PatchState& ps = lv.patches[p];
        for (int d = 0; d < 3; ++d)
        {
            g.patch[p].lower[d] = ps.lower[d];
            g.patch[p].upper[d] = ps.upper[d];
        }
        for (int a = 0; a < ndim; ++a)
        {
            const size_t cnt = (size_t)ps.n[a][0] * ps.n[a][1] * ps.n[a][2];
            const int* src = h_dof_index[p * ndim + a];
            if (!src) return fail(ctx, IBK_ERR_INVALID, "null DOF index array");
            CK(cudaMemcpyAsync(d_dof + off, src, sizeof(int) * cnt, cudaMemcpyHostToDevice, ctx->L.stream));
            g.patch[p].dof[a] = d_dof + off;
            for (int d = 0; d < 3; ++d) g.patch[p].n[a][d] = ps.n[a][d];
            off += cnt;
        }
    }
    int* d_cols = ctx->b_user[1].as<int>();
    int* d_count = d_cols + rows * npts;
    double* d_vals = ctx->b_user[2].as<double>();
    CK(cudaMemsetAsync(d_count, 0, sizeof(int), ctx->L.stream));
    const unsigned blocks = (unsigned)((rows + 127) / 128);
    if (S == 4)
        matop_rows_kernel<4><<<blocks, 128, 0, ctx->L.stream>>>(g, lv.X, lv.stride, lv.lag, n, d_cols, d_vals, d_count);
    else
        matop_rows_kernel<2><<<blocks, 128, 0, ctx->L.stream>>>(g, lv.X, lv.stride, lv.lag, n, d_cols, d_vals, d_count);
    ctx->L.launches++;
    CK(cudaGetLastError());
    int unplaced = 0;
    CK(cudaMemcpyAsync(h_cols, d_cols, sizeof(int) * rows * npts, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(h_vals, d_vals, sizeof(double) * rows * npts, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(&unplaced, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    if (n_unplaced) *n_unplaced = unplaced;
    return IBK_OK;
}
