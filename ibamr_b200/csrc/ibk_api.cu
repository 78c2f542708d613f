// ibk_api.cu -- the C ABI (include/ibk.h): context, raw funnel (seam B4), patch-level
// LEInteractor calls (seam B3).  The device-resident level (seams B1/B2) is in ibk_level.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ibk_ctx.h"

using namespace ibk;

namespace ibk
{
int fail(ibk_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    return code;
}
int cuda_fail(ibk_ctx* ctx, cudaError_t e, const char* what)
{
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return fail(ctx, IBK_ERR_CUDA, m);
}
int kernel_reach(int kernel)
{
    switch (kernel)
    {
    case IBK_PIECEWISE_LINEAR:
        return 1;
    case IBK_IB_6:
        return 3;
    default:
        return 2;
    }
}
long long round_pitch(int n0)
{
    return ((long long)n0 + 15) / 16 * 16; // rows start on 128-byte boundaries (TMA needs 16)
}
} // namespace ibk

#define CK(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// LEInteractor statics
// ---------------------------------------------------------------------------------------------
namespace ibk
{
int user_kernel_stencil_size(); // ibk_user.cu
}
extern "C" int ibk_kernel_from_string(const char* s)
{
    if (!s) return IBK_ERR_UNKNOWN_KERNEL;
    if (!strcmp(s, "PIECEWISE_LINEAR")) return IBK_PIECEWISE_LINEAR;
    if (!strcmp(s, "IB_4")) return IBK_IB_4;
    if (!strcmp(s, "IB_6")) return IBK_IB_6;
    if (!strcmp(s, "BSPLINE_3")) return IBK_BSPLINE_3;
    if (!strcmp(s, "BSPLINE_4")) return IBK_BSPLINE_4;
    if (!strcmp(s, "IB_3")) return IBK_IB_3;
    if (!strcmp(s, "BSPLINE_5")) return IBK_BSPLINE_5;
    if (!strcmp(s, "BSPLINE_6")) return IBK_BSPLINE_6;
    if (!strcmp(s, "PIECEWISE_CUBIC")) return IBK_PIECEWISE_CUBIC;
    if (!strcmp(s, "IB_5")) return IBK_IB_5;
    if (!strcmp(s, "PIECEWISE_CONSTANT")) return IBK_PIECEWISE_CONSTANT;
    if (!strcmp(s, "COMPOSITE_BSPLINE_32")) return IBK_COMPOSITE_BSPLINE_32;
    if (!strcmp(s, "COMPOSITE_BSPLINE_23")) return IBK_COMPOSITE_BSPLINE_23;
    if (!strcmp(s, "COMPOSITE_BSPLINE_43")) return IBK_COMPOSITE_BSPLINE_43;
    if (!strcmp(s, "COMPOSITE_BSPLINE_34")) return IBK_COMPOSITE_BSPLINE_34;
    if (!strcmp(s, "COMPOSITE_BSPLINE_54")) return IBK_COMPOSITE_BSPLINE_54;
    if (!strcmp(s, "COMPOSITE_BSPLINE_45")) return IBK_COMPOSITE_BSPLINE_45;
    if (!strcmp(s, "COMPOSITE_BSPLINE_65")) return IBK_COMPOSITE_BSPLINE_65;
    if (!strcmp(s, "COMPOSITE_BSPLINE_56")) return IBK_COMPOSITE_BSPLINE_56;
    if (!strcmp(s, "DISCONTINUOUS_LINEAR")) return IBK_DISCONTINUOUS_LINEAR;
    if (!strcmp(s, "IB_4_W8")) return IBK_IB_4_W8;
    if (!strcmp(s, "USER_DEFINED")) return IBK_USER_DEFINED;
    return IBK_ERR_UNKNOWN_KERNEL;
}
extern "C" int ibk_is_known_kernel(const char* s)
{
    return ibk_kernel_from_string(s) >= 0 ? 1 : 0;
}
extern "C" int ibk_get_stencil_size(const char* s)
{
    switch (ibk_kernel_from_string(s))
    {
    case IBK_PIECEWISE_LINEAR:
        return 2;
    case IBK_IB_4:
        return 4;
    case IBK_IB_6:
        return 6;
    case IBK_BSPLINE_3:
        return 4; // sic, LEInteractor.cpp:2057-2058
    case IBK_BSPLINE_4:
        return 4;
    case IBK_IB_3: // LEInteractor.cpp:2052-2103
        return 4;
    case IBK_BSPLINE_5:
        return 6;
    case IBK_BSPLINE_6:
        return 6;
    case IBK_PIECEWISE_CUBIC:
        return 4;
    case IBK_IB_5:
        return 6;
    case IBK_PIECEWISE_CONSTANT:
        return 1;
    case IBK_COMPOSITE_BSPLINE_32:
    case IBK_COMPOSITE_BSPLINE_23:
    case IBK_COMPOSITE_BSPLINE_43:
    case IBK_COMPOSITE_BSPLINE_34:
        return 4;
    case IBK_COMPOSITE_BSPLINE_54:
    case IBK_COMPOSITE_BSPLINE_45:
        return 5;
    case IBK_COMPOSITE_BSPLINE_65:
    case IBK_COMPOSITE_BSPLINE_56:
        return 6;
    case IBK_DISCONTINUOUS_LINEAR:
        return 2;
    case IBK_IB_4_W8:
        return 8;
    case IBK_USER_DEFINED:
        return ibk::user_kernel_stencil_size(); // s_kernel_fcn_stencil_size, LEInteractor.cpp:2099-2100
    default:
        return IBK_ERR_UNKNOWN_KERNEL;
    }
}
extern "C" int ibk_get_minimum_ghost_width(const char* s)
{
    const int sz = ibk_get_stencil_size(s);
    if (sz < 0) return sz;
    return (int)std::floor(0.5 * sz) + 1;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int ibk_ctx_create(int device, ibk_ctx** out)
{
    if (!out) return IBK_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    {
        cudaGetLastError();
        return IBK_ERR_CUDA; // no CPU fallback
    }
    if (cudaSetDevice(device) != cudaSuccess) return IBK_ERR_CUDA;
    ibk_ctx* ctx = new ibk_ctx();
    ctx->device = device;
    *out = ctx;
    return IBK_OK;
}

extern "C" int ibk_level_destroy(ibk_ctx* ctx);
extern "C" int ibk_comm_destroy(ibk_ctx* ctx);

extern "C" int ibk_ctx_destroy(ibk_ctx* ctx)
{
    if (!ctx) return IBK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->L.stream);
    ibk_comm_destroy(ctx);
    ibk_level_destroy(ctx);
    bins_free(ctx->sbins);
    ctx->b_Xe.release();
    ctx->b_Xr.release();
    ctx->b_Xes.release();
    ctx->b_Xrs.release();
    ctx->b_src.release();
    ctx->b_patchbin.release();
    for (auto& b : ctx->b_io) b.release();
    for (auto& b : ctx->b_stage) b.release();
    for (auto& b : ctx->b_mig) b.release();
    for (auto& b : ctx->b_user) b.release();
    if (ctx->ev_created)
        for (int i = 0; i < 3; ++i)
        {
            cudaEventDestroy(ctx->ev[i][0]);
            cudaEventDestroy(ctx->ev[i][1]);
        }
    if (ctx->xfer_created)
    {
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->s_out);
        for (int w = 0; w < 2; ++w)
        {
            cudaEventDestroy(ctx->ev_in[w]);
            cudaEventDestroy(ctx->ev_out[w]);
        }
        cudaEventDestroy(ctx->ev_order);
        cudaStreamDestroy(ctx->s_in);
        cudaStreamDestroy(ctx->s_out);
    }
    delete ctx;
    return IBK_OK;
}
extern "C" const char* ibk_last_error(const ibk_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : "null context";
}
extern "C" int ibk_ctx_set_stream(ibk_ctx* ctx, void* s)
{
    if (!ctx) return IBK_ERR_INVALID;
    ctx->L.stream = (cudaStream_t)s;
    return IBK_OK;
}
extern "C" int ibk_ctx_synchronize(ibk_ctx* ctx)
{
    if (!ctx) return IBK_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->L.stream));
    if (ctx->xfer_created)
    {
        CK(cudaStreamSynchronize(ctx->s_in));
        CK(cudaStreamSynchronize(ctx->s_out));
        for (int w = 0; w < 2; ++w) ctx->pend_in[w] = ctx->pend_out[w] = false;
    }
    return IBK_OK;
}
extern "C" long long ibk_ctx_launch_count(const ibk_ctx* ctx)
{
    return ctx ? ctx->L.launches : 0;
}
extern "C" int ibk_ctx_enable_timing(ibk_ctx* ctx, int enable)
{
    if (!ctx) return IBK_ERR_INVALID;
    if (enable && !ctx->ev_created)
    {
        for (int i = 0; i < 3; ++i)
        {
            CK(cudaEventCreate(&ctx->ev[i][0]));
            CK(cudaEventCreate(&ctx->ev[i][1]));
        }
        ctx->ev_created = true;
    }
    ctx->timing = enable != 0;
    return IBK_OK;
}
extern "C" int ibk_ctx_last_ms(ibk_ctx* ctx, int which, float* ms)
{
    if (!ctx || which < 0 || which > 2 || !ms) return IBK_ERR_INVALID;
    if (!ctx->ev_valid[which]) return fail(ctx, IBK_ERR_STATE, "no timing recorded");
    CK(cudaEventSynchronize(ctx->ev[which][1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// shared machinery of the raw / patch seams: entries -> bins -> tile kernel
// ---------------------------------------------------------------------------------------------
namespace ibk
{
struct ArrayComp
{
    double* ptr;
    long long pitch;
    int n[3];
    int nugc[3];
    int var[3];
    int vcol;
    int axis;
};

void make_tile_params(TileParams& tp, int ndim, const double* dx, const double xl[3][2], const int* nvar, const PatchBin& pb,
                      int ncomp, const ArrayComp* comps)
{
    std::memset(&tp, 0, sizeof(tp));
    tp.ndim = ndim;
    tp.ncomp = ncomp;
    tp.G = pb.G;
    tp.brick_base = pb.brick_base;
    double vol = 1.0;
    for (int d = 0; d < 3; ++d)
    {
        tp.dx[d] = d < ndim ? dx[d] : 1.0;
        tp.xl[d][0] = d < ndim ? xl[d][0] : 0.0;
        tp.xl[d][1] = d < ndim ? xl[d][1] : 0.0;
        tp.nvar[d] = d < ndim ? nvar[d] : 1;
        tp.nb[d] = pb.nb[d];
        tp.nt[d] = pb.nt[d];
        tp.ot_lo[d] = 0;
        tp.ot_n[d] = 1;
    }
    // dx(0)*dx(1)*dx(2) as the Fortran evaluates it (left to right), 3d.f.m4:1439
    vol = (ndim == 3) ? (dx[0] * dx[1]) * dx[2] : dx[0] * dx[1];
    tp.inv_vol = 1.0 / vol;
    for (int a = 0; a < ncomp; ++a)
    {
        CompGeom& c = tp.comp[a];
        c.ptr = comps[a].ptr;
        c.pitch = comps[a].pitch;
        c.vcol = comps[a].vcol;
        c.axis = comps[a].axis;
        for (int d = 0; d < 3; ++d)
        {
            c.n[d] = d < ndim ? comps[a].n[d] : 1;
            c.pp0[d] = d < ndim ? pb.G - comps[a].nugc[d] : 0;
            c.var[d] = d < ndim ? comps[a].var[d] : 0;
        }
    }
}

int user_kernel_stencil_size();
int user_entries_op(ibk_ctx* ctx, int op, const TileParams& tp, const CellGeom& cg, const PatchBin& pb, const double* d_Xe,
                    const double* d_Xr, long long stride, int n_entries, const int* d_indices, double* d_V, long long v_cstride,
                    long long v_istride, bool filter_box);
// Runs interp (op 0) or spread (op 1) for `n_entries` entries given as SoA positions Xe (shifted)
// and Xr (raw) of stride `stride`; values are addressed through d_indices (nullable).
int run_entries_op(ibk_ctx* ctx, int op, int kernel, TileParams& tp, const CellGeom& cg, PatchBin& pb, const double* d_Xe,
                   const double* d_Xr, long long stride, int n_entries, const int* d_indices, double* d_V, long long v_cstride,
                   long long v_istride, bool zero_unreached = true, const int* box_lo = nullptr, const int* box_hi = nullptr)
{
    if (n_entries <= 0) return IBK_OK;
    if (kernel == IBK_USER_DEFINED) // host callback: its own path (ibk_user.cu); position-only forms list by the box
        return user_entries_op(ctx, op, tp, cg, pb, d_Xe, d_Xr, stride, n_entries, d_indices, d_V, v_cstride, v_istride, !zero_unreached);
    const int ndim = tp.ndim;
    CK(ctx->b_patchbin.reserve(sizeof(PatchBin)));
    CK(cudaMemcpyAsync(ctx->b_patchbin.p, &pb, sizeof(PatchBin), cudaMemcpyHostToDevice, ctx->L.stream));
    CK(bins_build(ctx->sbins, ctx->L, cg, ctx->b_patchbin.as<PatchBin>(), 1, &pb, d_Xe, stride, nullptr,
                  (uint32_t)n_entries, n_entries, nullptr, nullptr));
    const uint32_t* perm = ctx->sbins.vals[ctx->sbins.sorted_in];
    CK(ctx->b_Xes.reserve(sizeof(double) * (size_t)ndim * n_entries));
    CK(gather_columns(ctx->L, d_Xe, stride, ctx->b_Xes.as<double>(), n_entries, perm, n_entries, ndim));
    const double* Xrs = nullptr;
    if (d_Xr)
    {
        CK(ctx->b_Xrs.reserve(sizeof(double) * (size_t)ndim * n_entries));
        CK(gather_columns(ctx->L, d_Xr, stride, ctx->b_Xrs.as<double>(), n_entries, perm, n_entries, ndim));
        Xrs = ctx->b_Xrs.as<double>();
    }
    CK(ctx->b_src.reserve(sizeof(uint32_t) * (size_t)n_entries));
    CK(compose_index(ctx->L, d_indices, perm, ctx->b_src.as<uint32_t>(), n_entries));
    MarkerView mv;
    mv.X = ctx->b_Xes.as<double>();
    mv.Xraw = Xrs;
    mv.x_stride = n_entries;
    mv.V = d_V;
    mv.v_cstride = v_cstride;
    mv.v_istride = v_istride;
    mv.src = ctx->b_src.as<uint32_t>();
    std::string err;
    // Index-list forms: a listed marker whose stencil finds no array point interpolates to 0 (the Fortran sets V = 0 and
    // adds nothing).  Position-only forms: the markers outside the box are simply not listed, their entries stay as
    // the caller left them (LEInteractor.cpp:6088-6126).
    if (op == 0 && zero_unreached)
        CK(zero_discarded(ctx->L, ctx->sbins.brick_start, ctx->sbins.total_bricks, n_entries, mv.src, d_V, v_cstride, v_istride,
                          tp.ncomp));
    else if (op == 0 && box_lo && box_hi) // position-only: the markers of the caller's box that the binning did not accept
        CK(zero_discarded_in_box(ctx->L, ctx->sbins.brick_start, ctx->sbins.total_bricks, n_entries, mv.X, mv.x_stride, cg, box_lo, box_hi,
                                 mv.src, d_V, v_cstride, v_istride, tp.ncomp));
    cudaError_t e = (op == 0) ? launch_interp(ctx->L, kernel, tp, ctx->sbins, mv, err) :
                                launch_spread(ctx->L, kernel, tp, ctx->sbins, mv, err);
    if (e != cudaSuccess) return cuda_fail(ctx, e, err.empty() ? "tile kernel launch" : err.c_str());
    return IBK_OK;
}
} // namespace ibk

// ---------------------------------------------------------------------------------------------
// seam B4: raw funnel
// ---------------------------------------------------------------------------------------------
static int raw_op(ibk_ctx* ctx, int op, int kernel, const ibk_array_desc* desc, double* d_u, long long pitch,
                  const int* d_indices, const double* d_Xshift, int nindices, const double* d_X, int n_markers, double* d_V)
{
    if (!ctx || !desc) return IBK_ERR_INVALID;
    if (kernel < 0 || (kernel > IBK_KERNEL_LAST && kernel != IBK_USER_DEFINED)) return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, "unknown kernel");
    const int ndim = desc->ndim;
    if (ndim != 2 && ndim != 3) return fail(ctx, IBK_ERR_INVALID, "ndim must be 2 or 3");
    if (desc->depth < 1 || desc->depth > IBK_MAX_COMP) return fail(ctx, IBK_ERR_INVALID, "depth out of range");
    if (nindices <= 0) return IBK_OK;
    (void)n_markers;
    int gmax = 0;
    ArrayComp comps[IBK_MAX_COMP];
    long long plane = 1;
    int n[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d)
    {
        n[d] = desc->iupper[d] - desc->ilower[d] + 1 + 2 * desc->nugc[d];
        gmax = std::max(gmax, desc->nugc[d]);
    }
    plane = pitch * n[1] * n[2];
    for (int c = 0; c < desc->depth; ++c)
    {
        comps[c].ptr = d_u + (size_t)c * plane;
        comps[c].pitch = pitch;
        comps[c].vcol = c;
        comps[c].axis = desc->axis;
        for (int d = 0; d < 3; ++d)
        {
            comps[c].n[d] = n[d];
            comps[c].nugc[d] = d < ndim ? desc->nugc[d] : 0;
            comps[c].var[d] = 0;
        }
    }
    const int G = gmax + 4;
    PatchBin pb;
    int alo[3], ahi[3];
    for (int d = 0; d < ndim; ++d)
    {
        alo[d] = desc->ilower[d] - G;
        ahi[d] = desc->iupper[d] + G;
    }
    fill_patch_bin(pb, ndim, desc->ilower, desc->iupper, alo, ahi, G, 0);
    CellGeom cg;
    std::memset(&cg, 0, sizeof(cg));
    cg.ndim = ndim;
    cg.two_branch = 0;
    double xl[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
    int nvar[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d)
    {
        cg.x_lower[d] = desc->x_lower[d];
        cg.x_upper[d] = desc->x_upper[d];
        cg.dx[d] = desc->dx[d];
        cg.ilower[d] = desc->ilower[d];
        cg.iupper[d] = desc->iupper[d];
        xl[d][0] = desc->x_lower[d];
    }
    TileParams tp;
    make_tile_params(tp, ndim, desc->dx, xl, nvar, pb, desc->depth, comps);
    CK(ctx->b_Xe.reserve(sizeof(double) * (size_t)ndim * nindices));
    CK(ctx->b_Xr.reserve(sizeof(double) * (size_t)ndim * nindices));
    CK(build_entries(ctx->L, d_X, d_indices, d_Xshift, nindices, ndim, ctx->b_Xe.as<double>(), ctx->b_Xr.as<double>(),
                     nindices));
    return run_entries_op(ctx, op, kernel, tp, cg, pb, ctx->b_Xe.as<double>(), d_Xshift ? ctx->b_Xr.as<double>() : nullptr,
                          nindices, nindices, d_indices, d_V, 1, desc->depth);
}

extern "C" int ibk_raw_interp(ibk_ctx* ctx, int kernel, const ibk_array_desc* desc, const double* d_u, const int* d_indices,
                              const double* d_Xshift, int nindices, const double* d_X, int n_markers, double* d_V)
{
    if (!desc) return IBK_ERR_INVALID;
    const int n0 = desc->iupper[0] - desc->ilower[0] + 1 + 2 * desc->nugc[0];
    return raw_op(ctx, 0, kernel, desc, const_cast<double*>(d_u), n0, d_indices, d_Xshift, nindices, d_X, n_markers, d_V);
}
extern "C" int ibk_raw_spread(ibk_ctx* ctx, int kernel, const ibk_array_desc* desc, const int* d_indices,
                              const double* d_Xshift, int nindices, const double* d_X, int n_markers, const double* d_V,
                              double* d_u)
{
    if (!desc) return IBK_ERR_INVALID;
    const int n0 = desc->iupper[0] - desc->ilower[0] + 1 + 2 * desc->nugc[0];
    return raw_op(ctx, 1, kernel, desc, d_u, n0, d_indices, d_Xshift, nindices, d_X, n_markers, const_cast<double*>(d_V));
}

static size_t desc_elems(const ibk_array_desc* desc, int* n)
{
    size_t t = 1;
    for (int d = 0; d < 3; ++d)
    {
        n[d] = d < desc->ndim ? desc->iupper[d] - desc->ilower[d] + 1 + 2 * desc->nugc[d] : 1;
        t *= (size_t)n[d];
    }
    return t;
}

static int raw_host(ibk_ctx* ctx, int op, int kernel, const ibk_array_desc* desc, double* h_u, const int* h_indices,
                    const double* h_Xshift, int nindices, const double* h_X, int n_markers, double* h_V)
{
    if (!ctx || !desc || !h_u || !h_X || !h_V || (nindices > 0 && !h_indices)) return IBK_ERR_INVALID;
    if (desc->ndim != 2 && desc->ndim != 3) return fail(ctx, IBK_ERR_INVALID, "ndim must be 2 or 3");
    if (desc->depth < 1 || desc->depth > IBK_MAX_COMP) return fail(ctx, IBK_ERR_INVALID, "depth out of range");
    if (nindices <= 0) return IBK_OK;
    const int ndim = desc->ndim, depth = desc->depth;
    int n[3];
    desc_elems(desc, n);
    const long long pitch = round_pitch(n[0]);
    const size_t rows = (size_t)n[1] * n[2] * depth;
    cudaStream_t st = ctx->L.stream;
    CK(ctx->b_io[0].reserve(sizeof(double) * (size_t)pitch * rows));
    CK(ctx->b_io[1].reserve(sizeof(int) * (size_t)nindices));
    CK(ctx->b_io[2].reserve(sizeof(double) * (size_t)nindices * ndim));
    CK(ctx->b_io[3].reserve(sizeof(double) * (size_t)n_markers * ndim));
    CK(ctx->b_io[4].reserve(sizeof(double) * (size_t)n_markers * depth));
    double* d_u = ctx->b_io[0].as<double>();
    // dense -> pitched, all depth slices at once (rows = n1*n2*depth)
    CK(cudaMemcpy2DAsync(d_u, (size_t)pitch * 8, h_u, (size_t)n[0] * 8, (size_t)n[0] * 8, rows, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_io[1].p, h_indices, sizeof(int) * (size_t)nindices, cudaMemcpyHostToDevice, st));
    if (h_Xshift)
        CK(cudaMemcpyAsync(ctx->b_io[2].p, h_Xshift, sizeof(double) * (size_t)nindices * ndim, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_io[3].p, h_X, sizeof(double) * (size_t)n_markers * ndim, cudaMemcpyHostToDevice, st));
    // V: interp must leave unlisted markers untouched -> upload the caller's values first
    CK(cudaMemcpyAsync(ctx->b_io[4].p, h_V, sizeof(double) * (size_t)n_markers * depth, cudaMemcpyHostToDevice, st));
    int rc = raw_op(ctx, op, kernel, desc, d_u, pitch, ctx->b_io[1].as<int>(), h_Xshift ? ctx->b_io[2].as<double>() : nullptr,
                    nindices, ctx->b_io[3].as<double>(), n_markers, ctx->b_io[4].as<double>());
    if (rc != IBK_OK) return rc;
    if (op == 0)
        CK(cudaMemcpyAsync(h_V, ctx->b_io[4].p, sizeof(double) * (size_t)n_markers * depth, cudaMemcpyDeviceToHost, st));
    else
        CK(cudaMemcpy2DAsync(h_u, (size_t)n[0] * 8, d_u, (size_t)pitch * 8, (size_t)n[0] * 8, rows, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return IBK_OK;
}

extern "C" int ibk_raw_interp_host(ibk_ctx* ctx, int kernel, const ibk_array_desc* desc, const double* h_u,
                                   const int* h_indices, const double* h_Xshift, int nindices, const double* h_X,
                                   int n_markers, double* h_V)
{
    return raw_host(ctx, 0, kernel, desc, const_cast<double*>(h_u), h_indices, h_Xshift, nindices, h_X, n_markers, h_V);
}
extern "C" int ibk_raw_spread_host(ibk_ctx* ctx, int kernel, const ibk_array_desc* desc, const int* h_indices,
                                   const double* h_Xshift, int nindices, const double* h_X, int n_markers, const double* h_V,
                                   double* h_u)
{
    return raw_host(ctx, 1, kernel, desc, h_u, h_indices, h_Xshift, nindices, h_X, n_markers, const_cast<double*>(h_V));
}

// ---------------------------------------------------------------------------------------------
// seam B3: patch-level LEInteractor calls on host data
// ---------------------------------------------------------------------------------------------
// centering: 0 = side (ndim arrays, one per axis, shifted along the axis), 1 = cell (one array with q_depth slices),
//            2 = node (one array with q_depth slices, shifted in every dimension; LEInteractor.cpp:2983-3043),
//            3 = edge (ndim arrays, array `axis` shifted in every dimension but `axis`; LEInteractor.cpp:3260-3340)
static int patch_host_op(ibk_ctx* ctx, int op, const char* fcn, const ibk_patch_desc* patch, int centering,
                         double* const* h_q, int q_depth, const int* box_lower, const int* box_upper,
                         const int* h_indices, const double* h_shifts, int n_indices, const double* h_X, int n_markers,
                         double* h_Q, int Q_depth)
{
    if (!ctx || !patch || !h_q || !h_X || !h_Q) return IBK_ERR_INVALID;
    const int kernel = ibk_kernel_from_string(fcn);
    if (kernel < 0) return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, std::string("unknown kernel function ") + (fcn ? fcn : "(null)"));
    const int ndim = patch->ndim;
    if (ndim != 2 && ndim != 3) return fail(ctx, IBK_ERR_INVALID, "ndim must be 2 or 3");
    const bool per_axis = centering == 0 || centering == 3; // one array per axis, vector-valued Lagrangian data
    if (per_axis && (Q_depth != ndim || q_depth != 1))
        return fail(ctx, IBK_ERR_DEPTH, centering == 0 ? "side-centered interpolation/spreading requires vector-valued data" :
                                                         "edge-centered interpolation/spreading requires vector-valued data");
    if (!per_axis && (Q_depth != q_depth || q_depth < 1 || q_depth > IBK_MAX_COMP))
        return fail(ctx, IBK_ERR_DEPTH, "Q_depth must equal the data depth");
    // is dimension d of array a shifted by half a cell (index i at x_lower + i dx instead of x_lower + (i + 1/2) dx)?
    auto shifted = [&](int a, int d) { return centering == 0 ? d == a : centering == 2 ? true : centering == 3 ? d != a : false; };
    // ghost-width validation: LEInteractor.cpp:4488-4498 (interp: always), :5250-5266 (spread: only
    // when the patch touches a physical boundary)
    const int min_ghosts = ibk_get_minimum_ghost_width(fcn);
    int gmin = patch->gcw[0], gmax = patch->gcw[0];
    for (int d = 1; d < ndim; ++d)
    {
        gmin = std::min(gmin, patch->gcw[d]);
        gmax = std::max(gmax, patch->gcw[d]);
    }
    if (gmin < min_ghosts && (op == 0 || patch->touches_physical_bdry))
    {
        char buf[256];
        snprintf(buf, sizeof(buf), "insufficient ghost cells: kernel function = %s, minimum ghost cell width = %d, ghost cell width = %d",
                 fcn, min_ghosts, gmin);
        return fail(ctx, IBK_ERR_GHOST_WIDTH, buf);
    }
    const bool indexed = h_indices != nullptr;
    const int n_entries = indexed ? n_indices : n_markers;
    if (n_entries <= 0) return IBK_OK;
    cudaStream_t st = ctx->L.stream;

    // ---- geometry
    const int ncomp = per_axis ? ndim : q_depth;
    ArrayComp comps[IBK_MAX_COMP];
    size_t off[IBK_MAX_COMP + 1];
    off[0] = 0;
    for (int a = 0; a < ncomp; ++a)
    {
        for (int d = 0; d < 3; ++d)
        {
            const int sh = (d < ndim && shifted(a, d)) ? 1 : 0;
            comps[a].n[d] = d < ndim ? patch->upper[d] - patch->lower[d] + 1 + 2 * patch->gcw[d] + sh : 1;
            comps[a].nugc[d] = d < ndim ? patch->gcw[d] : 0;
            comps[a].var[d] = sh;
        }
        comps[a].pitch = round_pitch(comps[a].n[0]);
        comps[a].vcol = a;
        comps[a].axis = per_axis ? a : 0; // LEInteractor passes the SideData / EdgeData axis, 0 for Cell / Node data
        off[a + 1] = off[a] + (size_t)comps[a].pitch * comps[a].n[1] * comps[a].n[2];
    }
    CK(ctx->b_io[0].reserve(sizeof(double) * off[ncomp]));
    for (int a = 0; a < ncomp; ++a)
    {
        comps[a].ptr = ctx->b_io[0].as<double>() + off[a];
        const double* src = per_axis ? h_q[a] : h_q[0] + (size_t)a * comps[a].n[0] * comps[a].n[1] * comps[a].n[2];
        CK(copy_dense_to_pitched(ctx->L, src, comps[a].ptr, comps[a].pitch, comps[a].n, ndim, cudaMemcpyHostToDevice));
    }
    const int G = gmax + 4;
    PatchBin pb;
    int alo[3], ahi[3];
    for (int d = 0; d < ndim; ++d)
    {
        alo[d] = patch->lower[d] - G;
        ahi[d] = patch->upper[d] + G;
        if (!indexed && box_lower && box_upper)
        {
            alo[d] = std::max(alo[d], box_lower[d]);
            ahi[d] = std::min(ahi[d], box_upper[d]);
        }
    }
    fill_patch_bin(pb, ndim, patch->lower, patch->upper, alo, ahi, G, 0);
    CellGeom cg;
    std::memset(&cg, 0, sizeof(cg));
    cg.ndim = ndim;
    cg.two_branch = 1; // IndexUtilities::getCellIndex(X, patch_geom, patch_box), LEInteractor.cpp:6113-6118
    double xl[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
    int nvar[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d)
    {
        cg.x_lower[d] = patch->x_lower[d];
        cg.x_upper[d] = patch->x_upper[d];
        cg.dx[d] = patch->dx[d];
        cg.ilower[d] = patch->lower[d];
        cg.iupper[d] = patch->upper[d];
        xl[d][0] = patch->x_lower[d];
        xl[d][1] = patch->x_lower[d] - 0.5 * patch->dx[d]; // x_lower_axis[axis] -= 0.5 * dx[axis], LEInteractor.cpp:2464
        nvar[d] = centering == 1 ? 1 : 2;
    }
    TileParams tp;
    make_tile_params(tp, ndim, patch->dx, xl, nvar, pb, ncomp, comps);

    // ---- marker data
    CK(ctx->b_io[3].reserve(sizeof(double) * (size_t)n_markers * ndim));
    CK(ctx->b_io[4].reserve(sizeof(double) * (size_t)n_markers * Q_depth));
    CK(cudaMemcpyAsync(ctx->b_io[3].p, h_X, sizeof(double) * (size_t)n_markers * ndim, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_io[4].p, h_Q, sizeof(double) * (size_t)n_markers * Q_depth, cudaMemcpyHostToDevice, st));
    const int* d_idx = nullptr;
    const double* d_shift = nullptr;
    if (indexed)
    {
        CK(ctx->b_io[1].reserve(sizeof(int) * (size_t)n_entries));
        CK(cudaMemcpyAsync(ctx->b_io[1].p, h_indices, sizeof(int) * (size_t)n_entries, cudaMemcpyHostToDevice, st));
        d_idx = ctx->b_io[1].as<int>();
        if (h_shifts)
        {
            CK(ctx->b_io[2].reserve(sizeof(double) * (size_t)n_entries * ndim));
            CK(cudaMemcpyAsync(ctx->b_io[2].p, h_shifts, sizeof(double) * (size_t)n_entries * ndim, cudaMemcpyHostToDevice, st));
            d_shift = ctx->b_io[2].as<double>();
        }
    }
    CK(ctx->b_Xe.reserve(sizeof(double) * (size_t)ndim * n_entries));
    CK(ctx->b_Xr.reserve(sizeof(double) * (size_t)ndim * n_entries));
    CK(build_entries(ctx->L, ctx->b_io[3].as<double>(), d_idx, d_shift, n_entries, ndim, ctx->b_Xe.as<double>(),
                     ctx->b_Xr.as<double>(), n_entries));
    int rc = run_entries_op(ctx, op, kernel, tp, cg, pb, ctx->b_Xe.as<double>(), d_shift ? ctx->b_Xr.as<double>() : nullptr,
                            n_entries, n_entries, d_idx, ctx->b_io[4].as<double>(), 1, Q_depth, /*zero_unreached*/ indexed,
                            indexed ? nullptr : box_lower, indexed ? nullptr : box_upper);
    if (rc != IBK_OK) return rc;
    if (op == 0)
    {
        CK(cudaMemcpyAsync(h_Q, ctx->b_io[4].p, sizeof(double) * (size_t)n_markers * Q_depth, cudaMemcpyDeviceToHost, st));
    }
    else
    {
        for (int a = 0; a < ncomp; ++a)
        {
            double* dst = per_axis ? h_q[a] : h_q[0] + (size_t)a * comps[a].n[0] * comps[a].n[1] * comps[a].n[2];
            CK(copy_pitched_to_dense(ctx->L, comps[a].ptr, comps[a].pitch, dst, comps[a].n, ndim, cudaMemcpyDeviceToHost));
        }
    }
    CK(cudaStreamSynchronize(st));
    return IBK_OK;
}

extern "C" int ibk_side_interpolate_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, const double* const* h_q,
                                         int q_depth, const int* box_lower, const int* box_upper, const double* h_X, int X_size,
                                         int X_depth, double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 0, fcn, patch, 0, const_cast<double* const*>(h_q), q_depth, box_lower, box_upper, nullptr,
                         nullptr, 0, h_X, X_size / X_depth, h_Q, Q_depth);
}
extern "C" int ibk_side_spread_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, double* const* h_q, int q_depth,
                                    const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth,
                                    const double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 1, fcn, patch, 0, h_q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         const_cast<double*>(h_Q), Q_depth);
}
extern "C" int ibk_cell_interpolate_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, const double* h_q,
                                         int q_depth, const int* box_lower, const int* box_upper, const double* h_X, int X_size,
                                         int X_depth, double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    double* q = const_cast<double*>(h_q);
    return patch_host_op(ctx, 0, fcn, patch, 1, &q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         h_Q, Q_depth);
}
extern "C" int ibk_cell_spread_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, double* h_q, int q_depth,
                                    const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth,
                                    const double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 1, fcn, patch, 1, &h_q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         const_cast<double*>(h_Q), Q_depth);
}
// NodeData (one array, any depth) and EdgeData (one array per axis, vector-valued), position-only overloads
extern "C" int ibk_node_interpolate_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, const double* h_q,
                                         int q_depth, const int* box_lower, const int* box_upper, const double* h_X, int X_size,
                                         int X_depth, double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    double* q = const_cast<double*>(h_q);
    return patch_host_op(ctx, 0, fcn, patch, 2, &q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         h_Q, Q_depth);
}
extern "C" int ibk_node_spread_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, double* h_q, int q_depth,
                                    const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth,
                                    const double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 1, fcn, patch, 2, &h_q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         const_cast<double*>(h_Q), Q_depth);
}
extern "C" int ibk_edge_interpolate_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, const double* const* h_q,
                                         int q_depth, const int* box_lower, const int* box_upper, const double* h_X, int X_size,
                                         int X_depth, double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 0, fcn, patch, 3, const_cast<double* const*>(h_q), q_depth, box_lower, box_upper, nullptr,
                         nullptr, 0, h_X, X_size / X_depth, h_Q, Q_depth);
}
extern "C" int ibk_edge_spread_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, double* const* h_q, int q_depth,
                                    const int* box_lower, const int* box_upper, const double* h_X, int X_size, int X_depth,
                                    const double* h_Q, int Q_size, int Q_depth)
{
    if (!patch) return IBK_ERR_INVALID;
    if (X_depth != patch->ndim) return fail(ctx, IBK_ERR_INVALID, "X_depth must be NDIM");
    (void)Q_size;
    return patch_host_op(ctx, 1, fcn, patch, 3, h_q, q_depth, box_lower, box_upper, nullptr, nullptr, 0, h_X, X_size / X_depth,
                         const_cast<double*>(h_Q), Q_depth);
}
extern "C" int ibk_side_interpolate_indexed_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch,
                                                 const double* const* h_q, const int* h_local_indices,
                                                 const double* h_periodic_shifts, int n_indices, const double* h_X,
                                                 int n_markers, double* h_Q)
{
    if (!patch || !h_local_indices) return IBK_ERR_INVALID;
    return patch_host_op(ctx, 0, fcn, patch, 0, const_cast<double* const*>(h_q), 1, nullptr, nullptr, h_local_indices,
                         h_periodic_shifts, n_indices, h_X, n_markers, h_Q, patch->ndim);
}
extern "C" int ibk_side_spread_indexed_host(ibk_ctx* ctx, const char* fcn, const ibk_patch_desc* patch, double* const* h_q,
                                            const int* h_local_indices, const double* h_periodic_shifts, int n_indices,
                                            const double* h_X, int n_markers, const double* h_Q)
{
    if (!patch || !h_local_indices) return IBK_ERR_INVALID;
    return patch_host_op(ctx, 1, fcn, patch, 0, h_q, 1, nullptr, nullptr, h_local_indices, h_periodic_shifts, n_indices, h_X,
                         n_markers, const_cast<double*>(h_Q), patch->ndim);
}
