// ibk_user.cu -- the USER_DEFINED kernel function of LEInteractor (row N4 of SURVEY 8(f)).
//
// Reference: the statics LEInteractor::s_kernel_fcn / s_kernel_fcn_stencil_size (ibtk/include/ibtk/LEInteractor.h:82-83,
// defaults LEInteractor.cpp:2019-2020: the 4-point function, stencil 4) and the two routines the funnel switches to for
// "USER_DEFINED" (LEInteractor.cpp:5207, 5981): userDefinedInterpolate (:6128-6257) and userDefinedSpread (:6259-6382).
//
// The kernel is a HOST callback, so its values can only be produced on the host: for every listed entry, component and
// dimension the host works out the clamped stencil range and calls the function for each of its points, exactly as the
// reference does (cell by floor, cell centre, left / right choice on the UNSHIFTED position for even stencils, clamp to the
// ghost box).  Everything that touches grid data runs on the device:
//   interpolation: one thread per (entry, component) sums w0 w1 w2 q over its range in the reference's loop order;
//   spreading:     every contribution (array element, list position, value) is emitted, the triples are sorted by
//                  (element, list position) with the library's radix sort and every element adds its contributions in list
//                  order: the reference's serial order of additions at every grid point, bit for bit, without atomics.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
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

namespace
{
// the reference's default for s_kernel_fcn: the 4-point function (ib4_kernel_fcn, LEInteractor.cpp:1526-1546), same operations
double default_kernel_fcn(double r)
{
    const double a = std::abs(r);
    if (a >= 2.0) return 0.0;
    const double a2 = a * a;
    if (a < 1.0) return -a / 4.0 + 3.0 / 8.0 + std::sqrt(-4.0 * a2 + 4.0 * a + 1.0) / 8.0;
    return -a / 4.0 + 5.0 / 8.0 - std::sqrt(12.0 * a - 7.0 - 4.0 * a2) / 8.0;
}
ibk_kernel_fcn g_user_fcn = &default_kernel_fcn;
int g_user_stencil = 4;
} // namespace

int user_kernel_stencil_size()
{
    return g_user_stencil;
}

constexpr int USER_MAX_STENCIL = 12;

struct UserComp
{
    double* ptr;
    long long pitch;
    int n[3], nug[3];
    int vcol;
    long long elem0; // number of this component's first element in the concatenation of all components (sort key)
};
struct UserGeom
{
    int ndim, ncomp, S;
    UserComp comp[IBK_MAX_COMP];
    double vol; // (dx0 * dx1) * dx2
};

// lo[(l * ncomp + a) * 3 + d], cnt[...]: first array coordinate and number of points; w[((l * ncomp + a) * 3 + d) * S + j]
__global__ void user_interp_kernel(const __grid_constant__ UserGeom g, const int* __restrict__ lo, const int* __restrict__ cnt,
                                   const double* __restrict__ w, const int* __restrict__ rows, int n_entries, double* __restrict__ V,
                                   long long v_cstride, long long v_istride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_entries * g.ncomp) return;
    const int l = t / g.ncomp, a = t % g.ncomp;
    const UserComp& c = g.comp[a];
    const int* L = lo + (size_t)t * 3;
    const int* C = cnt + (size_t)t * 3;
    const double* W = w + (size_t)t * 3 * g.S;
    double q = 0.0;
    const int c2 = g.ndim == 3 ? C[2] : 1;
    for (int k2 = 0; k2 < c2; ++k2)
        for (int k1 = 0; k1 < C[1]; ++k1)
            for (int k0 = 0; k0 < C[0]; ++k0)
            {
                double ww = __dmul_rn(W[k0], W[g.S + k1]);
                if (g.ndim == 3) ww = __dmul_rn(ww, W[2 * g.S + k2]);
                const long long e = ((long long)(g.ndim == 3 ? L[2] + k2 : 0) * c.n[1] + (L[1] + k1)) * c.pitch + (L[0] + k0);
                q = __dadd_rn(q, __dmul_rn(ww, c.ptr[e]));
            }
    const long long row = rows ? rows[l] : l;
    V[c.vcol * v_cstride + row * v_istride] = q;
}

// one thread per (entry, component): its contributions, S^ndim slots each (unused slots of a clamped stencil get the key ~0)
__global__ void user_pairs_kernel(const __grid_constant__ UserGeom g, const int* __restrict__ lo, const int* __restrict__ cnt,
                                  const double* __restrict__ w, const int* __restrict__ rows, int n_entries, const double* __restrict__ V,
                                  long long v_cstride, long long v_istride, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                  double* __restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_entries * g.ncomp) return;
    const int l = t / g.ncomp, a = t % g.ncomp;
    const UserComp& c = g.comp[a];
    const int* L = lo + (size_t)t * 3;
    const int* C = cnt + (size_t)t * 3;
    const double* W = w + (size_t)t * 3 * g.S;
    const long long row = rows ? rows[l] : l;
    const double Q = V[c.vcol * v_cstride + row * v_istride];
    int npts = 1;
    for (int d = 0; d < g.ndim; ++d) npts *= g.S;
    const size_t base = (size_t)t * npts;
    int s = 0;
    const int c2 = g.ndim == 3 ? C[2] : 1;
    for (int k2 = 0; k2 < c2; ++k2)
        for (int k1 = 0; k1 < C[1]; ++k1)
            for (int k0 = 0; k0 < C[0]; ++k0, ++s)
            {
                double ww = __dmul_rn(W[k0], W[g.S + k1]);
                if (g.ndim == 3) ww = __dmul_rn(ww, W[2 * g.S + k2]);
                // dense (unpitched) element number inside the component: the key only has to identify the element
                const long long e = ((long long)(g.ndim == 3 ? L[2] + k2 : 0) * c.n[1] + (L[1] + k1)) * c.n[0] + (L[0] + k0);
                keys[base + s] = ((uint64_t)(c.elem0 + e) << 32) | (uint32_t)l;
                idx[base + s] = (uint32_t)(base + s);
                vals[base + s] = __ddiv_rn(__dmul_rn(ww, Q), g.vol);
            }
    for (; s < npts; ++s)
    {
        keys[base + s] = ~0ull;
        idx[base + s] = (uint32_t)(base + s);
        vals[base + s] = 0.0;
    }
}

// the first pair of every element adds the element's pairs, in sorted (= list) order, to the array
__global__ void user_segsum_kernel(const __grid_constant__ UserGeom g, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                   const double* __restrict__ vals, long long n_pairs)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const uint64_t k = keys[i];
    if (k == ~0ull) return;
    const uint32_t elem = (uint32_t)(k >> 32);
    if (i > 0 && (uint32_t)(keys[i - 1] >> 32) == elem) return;
    int a = 0;
    while (a + 1 < g.ncomp && (long long)elem >= g.comp[a + 1].elem0) ++a;
    const UserComp& c = g.comp[a];
    long long e = (long long)elem - c.elem0;
    const int j0 = (int)(e % c.n[0]);
    e /= c.n[0];
    const int j1 = (int)(e % c.n[1]), j2 = (int)(e / c.n[1]);
    double* p = c.ptr + ((long long)j2 * c.n[1] + j1) * c.pitch + j0;
    double acc = *p;
    for (long long j = i; j < n_pairs && (uint32_t)(keys[j] >> 32) == elem && keys[j] != ~0ull; ++j) acc = __dadd_rn(acc, vals[idx[j]]);
    *p = acc;
}

#define CK(call)                                                   \
    do                                                             \
    {                                                              \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// Entries as run_entries_op receives them (ibk_api.cu): d_Xe shifted / d_Xr raw positions, SoA with `stride`; values through
// d_indices (nullable).  With filter_box the entries whose cell (cg) lies outside pb's accept box are not listed
// (LEInteractor::buildLocalIndices for the position-only forms); with index lists every entry is processed, as the funnel does.
int user_entries_op(ibk_ctx* ctx, int op, const TileParams& tp, const CellGeom& cg, const PatchBin& pb, const double* d_Xe,
                    const double* d_Xr, long long stride, int n_entries, const int* d_indices, double* d_V, long long v_cstride,
                    long long v_istride, bool filter_box)
{
    if (n_entries <= 0) return IBK_OK;
    const int ndim = tp.ndim, ncomp = tp.ncomp, S = g_user_stencil;
    cudaStream_t st = ctx->L.stream;
    std::vector<double> Xe((size_t)ndim * n_entries), Xr;
    CK(cudaMemcpy2DAsync(Xe.data(), sizeof(double) * n_entries, d_Xe, sizeof(double) * stride, sizeof(double) * n_entries, ndim,
                         cudaMemcpyDeviceToHost, st));
    if (d_Xr)
    {
        Xr.resize(Xe.size());
        CK(cudaMemcpy2DAsync(Xr.data(), sizeof(double) * n_entries, d_Xr, sizeof(double) * stride, sizeof(double) * n_entries, ndim,
                             cudaMemcpyDeviceToHost, st));
    }
    std::vector<int> rows_all;
    if (d_indices)
    {
        rows_all.resize(n_entries);
        CK(cudaMemcpyAsync(rows_all.data(), d_indices, sizeof(int) * n_entries, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    const std::vector<double>& XR = d_Xr ? Xr : Xe;

    UserGeom g;
    std::memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    g.ncomp = ncomp;
    g.S = S;
    g.vol = tp.dx[0] * tp.dx[1];
    if (ndim == 3) g.vol = g.vol * tp.dx[2];
    long long elem0 = 0;
    for (int a = 0; a < ncomp; ++a)
    {
        UserComp& c = g.comp[a];
        c.ptr = tp.comp[a].ptr;
        c.pitch = tp.comp[a].pitch;
        c.vcol = tp.comp[a].vcol;
        for (int d = 0; d < 3; ++d)
        {
            c.n[d] = tp.comp[a].n[d];
            c.nug[d] = d < ndim ? tp.G - tp.comp[a].pp0[d] : 0;
        }
        c.elem0 = elem0;
        elem0 += (long long)c.n[0] * c.n[1] * c.n[2];
    }
    if (elem0 >= (1ll << 32) - 1) return fail(ctx, IBK_ERR_INVALID, "USER_DEFINED: the arrays hold more than 2^32 elements");

    // ---- the list and the weights (host: the kernel function is a host callback)
    std::vector<int> list;
    list.reserve(n_entries);
    for (int l = 0; l < n_entries; ++l)
    {
        bool keep = true;
        if (filter_box)
            for (int d = 0; d < ndim && keep; ++d)
            {
                const double x = Xe[(size_t)d * n_entries + l];
                int cell;
                const double dlo = x - cg.x_lower[d], dup = x - cg.x_upper[d];
                if (!cg.two_branch || std::abs(dlo) <= std::abs(dup)) cell = cg.ilower[d] + (int)std::floor(dlo / cg.dx[d]);
                else cell = cg.iupper[d] + (int)std::floor(dup / cg.dx[d]) + 1;
                keep = cell >= pb.accept_lo[d] && cell <= pb.accept_hi[d];
            }
        if (keep) list.push_back(l);
    }
    const int nl = (int)list.size();
    if (nl == 0) return IBK_OK;
    std::vector<int> lo((size_t)nl * ncomp * 3, 0), cnt((size_t)nl * ncomp * 3, 1), rows(nl);
    std::vector<double> w((size_t)nl * ncomp * 3 * S, 0.0);
    for (int q = 0; q < nl; ++q)
    {
        const int l = list[q];
        rows[q] = d_indices ? rows_all[l] : l;
        for (int a = 0; a < ncomp; ++a)
            for (int d = 0; d < ndim; ++d)
            {
                const double xs = Xe[(size_t)d * n_entries + l], xraw = XR[(size_t)d * n_entries + l];
                const double x_lower = tp.xl[d][tp.comp[a].var[d]], dx = tp.dx[d];
                const int nug = g.comp[a].nug[d], iupper = g.comp[a].n[d] - 2 * nug - 1; // (relative to ilower = 0)
                const int center = (int)std::floor((xs - x_lower) / dx);
                const double x_cell = x_lower + ((double)center + 0.5) * dx;
                int slo, shi;
                if (S % 2 == 0)
                {
                    if (xraw < x_cell)
                    {
                        slo = center - S / 2;
                        shi = center + S / 2 - 1;
                    }
                    else
                    {
                        slo = center - S / 2 + 1;
                        shi = center + S / 2;
                    }
                }
                else
                {
                    slo = center - S / 2;
                    shi = center + S / 2;
                }
                slo = std::min(std::max(slo, -nug), iupper + nug);
                shi = std::min(std::max(shi, -nug), iupper + nug);
                const size_t o = ((size_t)q * ncomp + a) * 3 + d;
                lo[o] = slo + nug;
                cnt[o] = shi - slo + 1;
                for (int ic = slo; ic <= shi; ++ic) w[o * S + (ic - slo)] = g_user_fcn((xs - (x_cell + (double)(ic - center) * dx)) / dx);
            }
    }
    // ---- device
    DevBuf &b_lo = ctx->b_user[0], &b_cnt = ctx->b_user[1], &b_w = ctx->b_user[2], &b_rows = ctx->b_user[3];
    CK(b_lo.reserve(sizeof(int) * lo.size()));
    CK(b_cnt.reserve(sizeof(int) * cnt.size()));
    CK(b_w.reserve(sizeof(double) * w.size()));
    CK(b_rows.reserve(sizeof(int) * rows.size()));
    CK(cudaMemcpyAsync(b_lo.p, lo.data(), sizeof(int) * lo.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_cnt.p, cnt.data(), sizeof(int) * cnt.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_w.p, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_rows.p, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)(((size_t)nl * ncomp + 127) / 128);
    if (op == 0)
    {
        user_interp_kernel<<<blocks, 128, 0, st>>>(g, b_lo.as<int>(), b_cnt.as<int>(), b_w.as<double>(), b_rows.as<int>(), nl, d_V, v_cstride,
                                                   v_istride);
        ctx->L.launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st)); // (the host vectors above are the sources of asynchronous copies)
        return IBK_OK;
    }
    long long npts = 1;
    for (int d = 0; d < ndim; ++d) npts *= S;
    const long long n_pairs = (long long)nl * ncomp * npts;
    if (n_pairs > (1ll << 30)) return fail(ctx, IBK_ERR_INVALID, "USER_DEFINED spread: too many contributions for one call (split the index list)");
    DevBuf &b_ka = ctx->b_user[4], &b_kb = ctx->b_user[5], &b_va = ctx->b_user[6], &b_vb = ctx->b_user[7], &b_val = ctx->b_user[8], &b_tmp = ctx->b_user[9];
    CK(b_ka.reserve(sizeof(uint64_t) * n_pairs));
    CK(b_kb.reserve(sizeof(uint64_t) * n_pairs));
    CK(b_va.reserve(sizeof(uint32_t) * n_pairs));
    CK(b_vb.reserve(sizeof(uint32_t) * n_pairs));
    CK(b_val.reserve(sizeof(double) * n_pairs));
    CK(b_tmp.reserve(radix_sort_temp_bytes((int)n_pairs)));
    user_pairs_kernel<<<blocks, 128, 0, st>>>(g, b_lo.as<int>(), b_cnt.as<int>(), b_w.as<double>(), b_rows.as<int>(), nl, d_V, v_cstride, v_istride,
                                              b_ka.as<uint64_t>(), b_va.as<uint32_t>(), b_val.as<double>());
    ctx->L.launches++;
    const int which = radix_sort_pairs(b_ka.as<uint64_t>(), b_va.as<uint32_t>(), b_kb.as<uint64_t>(), b_vb.as<uint32_t>(), (int)n_pairs, 0, 64,
                                       b_tmp.p, st, &ctx->L.launches);
    const unsigned sblocks = (unsigned)((n_pairs + 255) / 256);
    user_segsum_kernel<<<sblocks, 256, 0, st>>>(g, which ? b_kb.as<uint64_t>() : b_ka.as<uint64_t>(), which ? b_vb.as<uint32_t>() : b_va.as<uint32_t>(),
                                                b_val.as<double>(), n_pairs);
    ctx->L.launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return IBK_OK;
}
} // namespace ibk

// LEInteractor::s_kernel_fcn / s_kernel_fcn_stencil_size: process-wide, like the reference's statics.  A null function
// restores the default (the 4-point function, stencil 4).
extern "C" int ibk_set_user_kernel(ibk_kernel_fcn fcn, int stencil_size)
{
    if (!fcn)
    {
        ibk::g_user_fcn = &ibk::default_kernel_fcn;
        ibk::g_user_stencil = 4;
        return IBK_OK;
    }
    if (stencil_size < 1 || stencil_size > ibk::USER_MAX_STENCIL) return IBK_ERR_INVALID;
    ibk::g_user_fcn = fcn;
    ibk::g_user_stencil = stencil_size;
    return IBK_OK;
}
