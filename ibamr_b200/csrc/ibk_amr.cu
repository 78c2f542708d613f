// ibk_amr.cu -- row N3 of SURVEY 8(f): the AMR transfer operators either side of the path, on side-centred data, between two
// levels that are registered on the same device (one context per level).
//
//  * prolongation of f before the spread on a finer level (f_prolongation_scheds[ln]->fillData, LDataManager.cpp:611-614;
//    the algorithm registers "CONSERVATIVE_LINEAR_REFINE", src/IB/IBHierarchyIntegrator.cpp:374-377);
//  * synchronisation of u before the interpolation, finest level first (f_synch_scheds[ln]->coarsenData,
//    LDataManager.cpp:728-734; "CONSERVATIVE_COARSEN", IBHierarchyIntegrator.cpp:369-372).
//
// Both operators are SAMRAI's (third party, IBSAMRAI2 >= 2025.10.29, not under /root/reference): for SideVariable<double>
// they are CartesianSideDoubleConservativeLinearRefine and CartesianSideDoubleWeightedAverage (source/geometry/cartesian/
// operators/side, Fortran cartclinrefsidedoub{2,3}d{0,1,2} and cartwgtavgsidedoub{2,3}d{0,1,2}).  Restated from their
// published algorithm (parity UNPINNED: no fixture of the reference holds their output):
//   refine:  fine(i) = c + sum_d slope_d * delta_d,  c = coarse(ic), ic = floor(i / ratio), ir = i - ic * ratio,
//            slope_d = monotonised central slope of the coarse data at ic in dimension d:
//                      dm = c - coarse(ic - e_d), dp = coarse(ic + e_d) - c, coef2 = (dm + dp) / 2, bound = 2 min(|dm|, |dp|),
//                      slope_d = dm * dp > 0 ? sign(min(|coef2|, bound), coef2) / dxc_d : 0,
//            delta_d = ir_d * dxf_d along the component's axis (a fine side on a coarse side takes its value),
//                      (ir_d + 1/2) dxf_d - dxc_d / 2 in the other dimensions (offset of the fine side's centre);
//   coarsen: coarse(ic) = (sum over the fine sides that tile the coarse side of fine * dAf) / dAc, dA = product of the mesh
//            widths of the dimensions other than the component's axis; the sum runs with the highest dimension outermost.
// Arithmetic is explicit round-to-nearest without contraction, so the oracle's numpy restatement is reproduced bit for bit.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "ibk_ctx.h"
#include "../../include/ibk.h"

namespace ibk
{
int fail(ibk_ctx* ctx, int code, const std::string& msg);
int cuda_fail(ibk_ctx* ctx, cudaError_t e, const char* what);

struct AmrArray
{
    double* p;
    long long pitch;
    int n1;     // rows per plane
    int lo[3];  // index of array element (0, 0, 0)
};
struct AmrJob
{
    AmrArray c, f;
    int ndim, axis;
    int lo[3], n[3]; // region: first index and extent (fine side indices for the refine, coarse ones for the coarsen)
    int ratio[3];
    double dxc[3], dxf[3];
};

__device__ __forceinline__ int floor_div(int a, int b)
{
    return a >= 0 ? a / b : -((-a + b - 1) / b);
}
__device__ __forceinline__ double& at(const AmrArray& A, const int (&i)[3])
{
    return A.p[((long long)(i[2] - A.lo[2]) * A.n1 + (i[1] - A.lo[1])) * A.pitch + (i[0] - A.lo[0])];
}

__global__ void amr_refine_side_kernel(const AmrJob j)
{
    const long long total = (long long)j.n[0] * j.n[1] * j.n[2];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
    {
        int i[3], ic[3], ir[3];
        i[0] = j.lo[0] + (int)(t % j.n[0]);
        i[1] = j.lo[1] + (int)((t / j.n[0]) % j.n[1]);
        i[2] = j.lo[2] + (int)(t / ((long long)j.n[0] * j.n[1]));
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            ic[d] = d < j.ndim ? floor_div(i[d], j.ratio[d]) : 0;
            ir[d] = d < j.ndim ? i[d] - ic[d] * j.ratio[d] : 0;
        }
        const double c = at(j.c, ic);
        double v = c;
        for (int d = 0; d < j.ndim; ++d)
        {
            int q[3] = { ic[0], ic[1], ic[2] };
            q[d] = ic[d] - 1;
            const double dm = __dsub_rn(c, at(j.c, q));
            q[d] = ic[d] + 1;
            const double dp = __dsub_rn(at(j.c, q), c);
            const double coef2 = __dmul_rn(0.5, __dadd_rn(dm, dp));
            const double bound = __dmul_rn(2.0, fmin(fabs(dm), fabs(dp)));
            double slope = 0.0;
            if (__dmul_rn(dm, dp) > 0.0) slope = __ddiv_rn(copysign(fmin(fabs(coef2), bound), coef2), j.dxc[d]);
            const double delta = d == j.axis ? __dmul_rn((double)ir[d], j.dxf[d]) :
                                               __dsub_rn(__dmul_rn(__dadd_rn((double)ir[d], 0.5), j.dxf[d]), __dmul_rn(j.dxc[d], 0.5));
            v = __dadd_rn(v, __dmul_rn(slope, delta));
        }
        at(j.f, i) = v;
    }
}

__global__ void amr_coarsen_side_kernel(const AmrJob j)
{
    const long long total = (long long)j.n[0] * j.n[1] * j.n[2];
    double dAf = 1.0, dAc = 1.0;
    for (int d = 0; d < j.ndim; ++d)
        if (d != j.axis)
        {
            dAf = __dmul_rn(dAf, j.dxf[d]);
            dAc = __dmul_rn(dAc, j.dxc[d]);
        }
    int r[3];
    for (int d = 0; d < 3; ++d) r[d] = (d < j.ndim && d != j.axis) ? j.ratio[d] : 1;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
    {
        int ic[3];
        ic[0] = j.lo[0] + (int)(t % j.n[0]);
        ic[1] = j.lo[1] + (int)((t / j.n[0]) % j.n[1]);
        ic[2] = j.lo[2] + (int)(t / ((long long)j.n[0] * j.n[1]));
        double s = 0.0;
        for (int k2 = 0; k2 < r[2]; ++k2)
            for (int k1 = 0; k1 < r[1]; ++k1)
                for (int k0 = 0; k0 < r[0]; ++k0)
                {
                    int q[3] = { 0, 0, 0 };
                    const int k[3] = { k0, k1, k2 };
                    for (int d = 0; d < j.ndim; ++d) q[d] = ic[d] * j.ratio[d] + (d == j.axis ? 0 : k[d]);
                    s = __dadd_rn(s, __dmul_rn(at(j.f, q), dAf));
                }
        at(j.c, ic) = __ddiv_rn(s, dAc);
    }
}
} // namespace ibk

using namespace ibk;

namespace
{
int host_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
int host_ceil_div(int a, int b) { return -host_floor_div(-a, b); }

AmrArray array_of(const LevelState& lv, const PatchState& ps, int which, int axis)
{
    AmrArray A;
    A.p = which == 0 ? ps.u[axis] : ps.f[axis];
    A.pitch = ps.pitch[axis];
    A.n1 = ps.n[axis][1];
    for (int d = 0; d < 3; ++d) A.lo[d] = d < lv.ndim ? ps.lower[d] - lv.gcw[d] : 0;
    return A;
}

// the destination context's stream waits for everything queued on the source context's stream
cudaError_t order_after(ibk_ctx* dst, ibk_ctx* src)
{
    if (dst->L.stream == src->L.stream) return cudaSuccess;
    cudaEvent_t ev;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev, src->L.stream)) == cudaSuccess) e = cudaStreamWaitEvent(dst->L.stream, ev, 0);
    cudaEventDestroy(ev); // (released when the recorded work has completed)
    return e;
}

int check_pair(ibk_ctx* coarse, ibk_ctx* fine, int which, const int* ratio, ibk_ctx* report)
{
    if (!coarse || !fine || !ratio) return IBK_ERR_INVALID;
    if (!coarse->lv.valid || !fine->lv.valid) return fail(report, IBK_ERR_STATE, "both contexts need a registered level (ibk_level_create)");
    if (coarse->device != fine->device) return fail(report, IBK_ERR_STATE, "the two levels must live on the same device");
    if (coarse->lv.ndim != fine->lv.ndim) return fail(report, IBK_ERR_INVALID, "the two levels differ in dimension");
    if (which != 0 && which != 1) return fail(report, IBK_ERR_INVALID, "which must be 0 (u) or 1 (f)");
    for (int d = 0; d < coarse->lv.ndim; ++d)
    {
        if (ratio[d] < 1) return fail(report, IBK_ERR_INVALID, "refinement ratio must be positive");
        const long long nc = coarse->lv.domain_upper[d] - coarse->lv.domain_lower[d] + 1;
        const long long nf = fine->lv.domain_upper[d] - fine->lv.domain_lower[d] + 1;
        if (nc * ratio[d] != nf || fine->lv.domain_lower[d] != coarse->lv.domain_lower[d] * ratio[d])
            return fail(report, IBK_ERR_INVALID, "the fine level's index space is not the coarse one refined by the ratio");
    }
    return IBK_OK;
}
} // namespace

#define CKC(ctx, call)                                                \
    do                                                                \
    {                                                                 \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call); \
    } while (0)

// fine `which` := CONSERVATIVE_LINEAR_REFINE(coarse `which`) at every point of the fine arrays (ghosts included) whose coarse
// stencil (the coarse point and its two neighbours per dimension) lies inside a coarse patch's array; coarse patches in
// list order (where their ghost regions overlap, the values agree once the coarse ghosts are filled: ibk_halo_local /
// the exchange first).  Returns in *n_points (may be null) how many fine points were written.
extern "C" int ibk_amr_refine_side(ibk_ctx* fine, ibk_ctx* coarse, int which, const int* ratio, long long* n_points)
{
    if (int rc = check_pair(coarse, fine, which, ratio, fine)) return rc;
    const LevelState& lc = coarse->lv;
    const LevelState& lf = fine->lv;
    CKC(fine, cudaSetDevice(fine->device));
    CKC(fine, order_after(fine, coarse));
    long long written = 0;
    for (const PatchState& pf : lf.patches)
        for (const PatchState& pc : lc.patches)
            for (int a = 0; a < lf.ndim; ++a)
            {
                AmrJob j;
                j.c = array_of(lc, pc, which, a);
                j.f = array_of(lf, pf, which, a);
                j.ndim = lf.ndim;
                j.axis = a;
                long long count = 1;
                for (int d = 0; d < 3; ++d)
                {
                    j.ratio[d] = d < lf.ndim ? ratio[d] : 1;
                    j.dxc[d] = lc.dx[d];
                    j.dxf[d] = lf.dx[d];
                    if (d >= lf.ndim)
                    {
                        j.lo[d] = 0;
                        j.n[d] = 1;
                        continue;
                    }
                    const int side = d == a ? 1 : 0;
                    // coarse centres whose +-1 neighbours are inside the coarse array
                    const int cu_lo = pc.lower[d] - lc.gcw[d] + 1, cu_hi = pc.upper[d] + lc.gcw[d] + side - 1;
                    const int flo = std::max(pf.lower[d] - lf.gcw[d], cu_lo * ratio[d]);
                    const int fhi = std::min(pf.upper[d] + lf.gcw[d] + side, cu_hi * ratio[d] + ratio[d] - 1);
                    j.lo[d] = flo;
                    j.n[d] = fhi - flo + 1;
                    count *= std::max(j.n[d], 0);
                }
                if (count <= 0) continue;
                const unsigned blocks = (unsigned)std::min<long long>((count + 255) / 256, 148 * 16);
                amr_refine_side_kernel<<<blocks, 256, 0, fine->L.stream>>>(j);
                fine->L.launches++;
                written += count;
            }
    CKC(fine, cudaGetLastError());
    CKC(coarse, order_after(coarse, fine)); // the coarse arrays may be rewritten once the refine has read them
    if (n_points) *n_points = written;
    return IBK_OK;
}

// coarse `which` := CONSERVATIVE_COARSEN(fine `which`) on the coarse patches' own sides (no ghosts) that are tiled by the
// own sides of a fine patch (the sides on the boundary of the refined region included).
extern "C" int ibk_amr_coarsen_side(ibk_ctx* coarse, ibk_ctx* fine, int which, const int* ratio, long long* n_points)
{
    if (int rc = check_pair(coarse, fine, which, ratio, coarse)) return rc;
    const LevelState& lc = coarse->lv;
    const LevelState& lf = fine->lv;
    CKC(coarse, cudaSetDevice(coarse->device));
    CKC(coarse, order_after(coarse, fine));
    long long written = 0;
    for (const PatchState& pc : lc.patches)
        for (const PatchState& pf : lf.patches)
            for (int a = 0; a < lc.ndim; ++a)
            {
                AmrJob j;
                j.c = array_of(lc, pc, which, a);
                j.f = array_of(lf, pf, which, a);
                j.ndim = lc.ndim;
                j.axis = a;
                long long count = 1;
                for (int d = 0; d < 3; ++d)
                {
                    j.ratio[d] = d < lc.ndim ? ratio[d] : 1;
                    j.dxc[d] = lc.dx[d];
                    j.dxf[d] = lf.dx[d];
                    if (d >= lc.ndim)
                    {
                        j.lo[d] = 0;
                        j.n[d] = 1;
                        continue;
                    }
                    int lo, hi;
                    if (d == a)
                    {
                        // coarse sides that coincide with a fine side of the patch's side box [lower, upper + 1]
                        lo = host_ceil_div(pf.lower[d], ratio[d]);
                        hi = host_floor_div(pf.upper[d] + 1, ratio[d]);
                        lo = std::max(lo, pc.lower[d]);
                        hi = std::min(hi, pc.upper[d] + 1);
                    }
                    else
                    {
                        // coarse cells completely covered by the fine patch
                        lo = host_ceil_div(pf.lower[d], ratio[d]);
                        hi = host_floor_div(pf.upper[d] + 1, ratio[d]) - 1;
                        lo = std::max(lo, pc.lower[d]);
                        hi = std::min(hi, pc.upper[d]);
                    }
                    j.lo[d] = lo;
                    j.n[d] = hi - lo + 1;
                    count *= std::max(j.n[d], 0);
                }
                if (count <= 0) continue;
                const unsigned blocks = (unsigned)std::min<long long>((count + 255) / 256, 148 * 16);
                amr_coarsen_side_kernel<<<blocks, 256, 0, coarse->L.stream>>>(j);
                coarse->L.launches++;
                written += count;
            }
    CKC(coarse, cudaGetLastError());
    CKC(fine, order_after(fine, coarse));
    if (n_points) *n_points = written;
    return IBK_OK;
}
