// ibk_level.cu -- the device-resident level behind seams B1/B2 (include/ibk.h): side-centred u / f
// for the patches this process owns, SoA fp64 marker columns (LData role), rebin
// (LDataManager::begin/endDataRedistribution role), spreadForce / interpolateVelocity
// (IBMethod.cpp:972-995 / :672-694 over LDataManager.cpp:551-667 / :698-813) and the halo ops
// that replace the SAMRAI schedules around them.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ibk_ctx.h"

using namespace ibk;

namespace ibk
{
int fail(ibk_ctx* ctx, int code, const std::string& msg);
int cuda_fail(ibk_ctx* ctx, cudaError_t e, const char* what);
int kernel_reach(int kernel);
long long round_pitch(int n0);
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
                      int ncomp, const ArrayComp* comps);

// ---------------------------------------------------------------------------------------------
// halo kernels
// ---------------------------------------------------------------------------------------------
constexpr int HALO_MAXSRC = 64;
struct HaloSrc
{
    const double* ptr;
    long long pitch;
    int n1;
    int alo[3], ahi[3]; // array box of the source, already shifted into the destination's index space
    int ilo[3], ihi[3]; // interior (side) box of the source, shifted likewise
};
struct HaloArgs
{
    double* dst;
    long long pitch;
    int n1;
    int ndim;
    int alo[3], ahi[3]; // array box of the destination
    int ilo[3], ihi[3]; // interior (side) box of the destination
    int rim;            // ghost width (thickness of the interior rim that ghosts of others can reach)
    int nsrc;
    HaloSrc src[HALO_MAXSRC];
};

// The local halo operations as lists of box operations worked out once per level (build_halo_plan):
//   MODE 0  dst box <- src box                      ghost fill of u
//   MODE 1  dst box += src_0 box + src_1 box + ...   ghost accumulation of f: the sources of a box in canonical order
//   MODE 2  dst box = 0                              ghost regions of f before a spread
// No per-element tests.  The CTAs of a launch are dealt out to the items in proportion to their size (block0 / nblocks),
// a warp walks a row (or a few short rows) of its item.
constexpr int REGION_MAXSRC = 8;
struct RegionMulti
{
    double* dst;
    long long dst_pitch;
    int dst_n1;
    int dst_off[3], ext[3];
    int nsrc;
    unsigned block0, nblocks;
    const double* src[REGION_MAXSRC];
    long long src_pitch[REGION_MAXSRC];
    int src_n1[REGION_MAXSRC];
    int src_off[REGION_MAXSRC][3];
};
template <int MODE>
__global__ void region_items_kernel(const RegionMulti* __restrict__ items, int n_items)
{
    int lo_i = 0, hi_i = n_items - 1; // the item this CTA belongs to: the last one with block0 <= blockIdx.x
    while (lo_i < hi_i)
    {
        const int mid = (lo_i + hi_i + 1) >> 1;
        if (items[mid].block0 <= blockIdx.x) lo_i = mid;
        else hi_i = mid - 1;
    }
    const RegionMulti& r = items[lo_i];
    const int lo[3] = { 0, 0, 0 }, hi[3] = { r.ext[0] - 1, r.ext[1] - 1, r.ext[2] - 1 };
    slab_rows(lo, hi, blockIdx.x - r.block0, r.nblocks, [&](int j, int k, int x0, int x1) {
        double* drow = r.dst + ((long long)(r.dst_off[2] + k) * r.dst_n1 + (r.dst_off[1] + j)) * r.dst_pitch + r.dst_off[0];
        if (MODE == 2)
        {
            for (int x = x0; x <= x1; x += 32) drow[x] = 0.0;
            return;
        }
        const double* srow[REGION_MAXSRC];
        for (int c = 0; c < r.nsrc; ++c)
            srow[c] = r.src[c] + ((long long)(r.src_off[c][2] + k) * r.src_n1[c] + (r.src_off[c][1] + j)) * r.src_pitch[c] + r.src_off[c][0];
        for (int x = x0; x <= x1; x += 32)
        {
            if (MODE == 0)
                drow[x] = srow[0][x];
            else
            {
                double acc = drow[x];
                for (int c = 0; c < r.nsrc; ++c) acc += srow[c][x];
                drow[x] = acc;
            }
        }
    });
}

struct FacePair
{
    double* a;
    double* b;
    long long a_pitch, b_pitch;
    int a_n1, b_n1;
    int a_off[3], b_off[3];
    int ext[3];
};
// Shared faces of the axis-normal component exist in two interiors (or twice in one periodic
// patch): both copies become a + b (PETScVecUtilities.cpp:519-611 gives them one DOF).
__global__ void face_sync_kernel(const FacePair* __restrict__ pairs)
{
    const FacePair& P = pairs[blockIdx.y];
    const long long total = (long long)P.ext[0] * P.ext[1] * P.ext[2];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q % P.ext[0]);
        const long long t = q / P.ext[0];
        const int j = (int)(t % P.ext[1]), k = (int)(t / P.ext[1]);
        double* pa = P.a + ((long long)(P.a_off[2] + k) * P.a_n1 + (P.a_off[1] + j)) * P.a_pitch + P.a_off[0] + i;
        double* pb = P.b + ((long long)(P.b_off[2] + k) * P.b_n1 + (P.b_off[1] + j)) * P.b_pitch + P.b_off[0] + i;
        const double s = *pa + *pb;
        *pa = s;
        *pb = s;
    }
}

// Pre-existing content of f on the two boundary face layers of the axis-normal component must not
// take part in the face/ghost sums (the reference spreads into a zeroed f and adds the old f to the
// interiors afterwards, LDataManager.cpp:589-594, 662-663): park it, zero the layers, restore-add later.
// mode 0: save + zero, mode 1: layer += saved.
struct FaceLayerJob
{
    const HaloArgs* args;
    int axis;
    double* save;
};
__global__ void face_layers_kernel(const FaceLayerJob* __restrict__ jobs, int mode)
{
    const HaloArgs& A = *jobs[blockIdx.y].args;
    const int axis = jobs[blockIdx.y].axis;
    double* __restrict__ save = jobs[blockIdx.y].save;
    int ext[3];
    for (int d = 0; d < 3; ++d) ext[d] = (d == axis) ? 1 : (A.ihi[d] - A.ilo[d] + 1);
    const long long per = (long long)ext[0] * ext[1] * ext[2];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < 2 * per; q += (long long)gridDim.x * blockDim.x)
    {
        const int side = (int)(q / per);
        long long r = q % per;
        int I[3];
        I[0] = A.ilo[0] + (int)(r % ext[0]);
        r /= ext[0];
        I[1] = A.ilo[1] + (int)(r % ext[1]);
        I[2] = A.ilo[2] + (int)(r / ext[1]);
        I[axis] = side ? A.ihi[axis] : A.ilo[axis];
        double* p = A.dst + ((long long)(I[2] - A.alo[2]) * A.n1 + (I[1] - A.alo[1])) * A.pitch + (I[0] - A.alo[0]);
        if (mode == 0)
        {
            save[q] = *p;
            *p = 0.0;
        }
        else
        {
            *p += save[q];
        }
    }
}

// Physical-boundary fold-back of the spread force: the adjoint of the linear ghost-cell extrapolation of a Robin boundary
// condition a u + b du/dn = g, CartSideRobinPhysBdryOp::accumulateFromPhysicalBoundaryData
// (ibtk/src/boundary/physical_boundary/CartSideRobinPhysBdryOp.cpp:552-617) with the co-dimension-one Fortran kernels in
// adjoint mode (fortran/cartphysbdryop3d.f.m4: scrobinphysbdryop1x3d :787-905 for the component normal to the wall,
// ccrobinphysbdryop1x3d :78-168 for the transverse ones; homogeneous form: g = 0).  One thread per column along the wall
// normal, over the WHOLE transverse extent of the array (ghost columns included: what a neighbouring patch's fold-back would
// have put into its own interior reaches it afterwards through the ghost accumulation).
struct WallJob
{
    double* ptr;
    long long stride[3]; // element strides of the array
    int n[3];            // extents incl. ghosts
    int dim, side;       // wall normal, 0 lower / 1 upper
    int normal;          // the component is the one normal to the wall
    int ib;              // normal: array index of the boundary face; transverse: of the first interior cell next to the wall
    int gcw;
    double a, b, h;
};
__global__ void wall_fold_kernel(const WallJob* __restrict__ jobs)
{
    const WallJob& J = jobs[blockIdx.y];
    const int d = J.dim, e1 = (d + 1) % 3, e2 = (d + 2) % 3;
    const long long total = (long long)J.n[e1] * J.n[e2];
    const int sgn = J.side ? +1 : -1;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int j1 = (int)(q % J.n[e1]), j2 = (int)(q / J.n[e1]);
        double* col = J.ptr + j1 * J.stride[e1] + j2 * J.stride[e2];
        if (J.normal)
        {
            const bool dirichlet = fabs(J.b) < 1.0e-12;
            double ub = dirichlet ? 0.0 : col[J.ib * J.stride[d]]; // (Dirichlet: u_b = g / a = 0 is written first, f.m4:864-865)
            for (int i = 1; i <= J.gcw; ++i)
            {
                const int ig = J.ib + sgn * i, ii = J.ib - sgn * i;
                if (ig < 0 || ig >= J.n[d] || ii < 0 || ii >= J.n[d]) continue;
                const double ug = col[ig * J.stride[d]];
                const double fi = dirichlet ? -1.0 : 1.0;
                const double fb = dirichlet ? 2.0 : -J.a * (2.0 * i) * J.h / J.b;
                col[ii * J.stride[d]] += fi * ug;
                ub += fb * ug;
            }
            col[J.ib * J.stride[d]] = ub;
        }
        else
        {
            for (int i = 0; i < J.gcw; ++i)
            {
                const int ig = J.ib + sgn * (1 + i), ii = J.ib - sgn * i;
                if (ig < 0 || ig >= J.n[d] || ii < 0 || ii >= J.n[d]) continue;
                const double nn = 1.0 + 2.0 * i;
                const double fi = -(J.a * nn * J.h - 2.0 * J.b) / (J.a * nn * J.h + 2.0 * J.b);
                col[ii * J.stride[d]] += fi * col[ig * J.stride[d]];
            }
        }
    }
}

__global__ void count_nonzero_kernel(const double* __restrict__ p, long long pitch, int n0, long long rows,
                                     unsigned long long* __restrict__ out)
{
    unsigned long long c = 0;
    const long long total = rows * n0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const long long r = q / n0;
        const int i = (int)(q % n0);
        if (p[r * pitch + i] != 0.0) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void iota_kernel(uint32_t* p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void gather_u32_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
} // namespace ibk

#define CK(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);                                                     \
    } while (0)
#define NEED_LEVEL()                                                                                                   \
    if (!ctx) return IBK_ERR_INVALID;                                                                                  \
    if (!ctx->lv.valid) return fail(ctx, IBK_ERR_STATE, "no level registered (ibk_level_create)")

// host-side halo plans live next to the level (file-local registry keyed by ctx)
namespace
{
struct HaloPlan
{
    // [which][patch][axis]
    std::vector<HaloArgs*> d_args[2];
    std::vector<FacePair*> d_pairs; // per axis (f only)
    std::vector<int> n_pairs;
    std::vector<double*> d_face_save; // [patch][axis]: 2 boundary layers of f's axis-normal component
    std::vector<void*> allocs;
    // box operations of the whole level: [0] ghost fill of u, [1] ghost accumulation of f (in waves: the items of a wave
    // write disjoint elements, the waves follow the canonical source order), [2] zero of the ghost regions of f
    RegionMulti* d_ops[3] = { nullptr, nullptr, nullptr };
    int n_ops[3] = { 0, 0, 0 };
    unsigned n_blocks[3] = { 0, 0, 0 };
    FacePair* d_all_pairs = nullptr; // the face pairs of all axes (one launch)
    int n_all_pairs = 0;
    WallJob* d_wall_jobs = nullptr; // physical-boundary fold-back (ibk_level_set_wall_bc)
    int n_wall_jobs = 0;
    unsigned wall_blocks = 1;
    bool walls_folded = false; // since the last ibk_spread_begin
    FaceLayerJob* d_face_jobs = nullptr; // [patch][axis]
    int n_face_jobs = 0;
    unsigned face_job_blocks = 1;
};
// device copy of the item table of one message (ibk_halo_pack_many / _unpack_many), keyed by its content
struct ItemTable
{
    unsigned long long key = 0;
    HaloItem* d_items = nullptr;
    int n = 0;
    // waves: maximal runs of consecutive items whose regions are pairwise disjoint; one launch adds a wave, the waves follow
    // each other in item order, so two overlapping regions are still added in the order of the list
    std::vector<HaloItem> h_items;    // host copy (for building the add table)
    std::vector<int> wave_start;      // [n_waves + 1]
    std::vector<unsigned> wave_blocks; // CTAs per wave
    unsigned all_blocks = 0;           // CTAs when all items go in one launch (pack)
    long long max_count = 0;
    // unpack-ADD in one launch: the destination regions cut into boxes that each have ONE ordered list of buffer sources
    // (built for the buffer address add_buf at the first unpack-add; region_items_kernel<1>)
    RegionMulti* d_add_ops = nullptr;
    int n_add_ops = 0;
    unsigned add_blocks = 0;
    const double* add_buf = nullptr;
};
struct LevelExtra
{
    HaloPlan halo;
    std::vector<TileParams> tp; // per patch: [u params, f params] interleaved: tp[2*p + which]
    std::vector<ItemTable> item_tables;
};
static std::vector<std::pair<ibk_ctx*, LevelExtra*>> g_extra;
LevelExtra* extra_of(ibk_ctx* ctx, bool create)
{
    for (auto& kv : g_extra)
        if (kv.first == ctx) return kv.second;
    if (!create) return nullptr;
    g_extra.push_back({ ctx, new LevelExtra() });
    return g_extra.back().second;
}
void extra_drop(ibk_ctx* ctx)
{
    for (size_t i = 0; i < g_extra.size(); ++i)
        if (g_extra[i].first == ctx)
        {
            for (void* p : g_extra[i].second->halo.allocs) cudaFree(p);
            if (g_extra[i].second->halo.d_wall_jobs) cudaFree(g_extra[i].second->halo.d_wall_jobs);
            for (ItemTable& t : g_extra[i].second->item_tables)
            {
                if (t.d_items) cudaFree(t.d_items);
                if (t.d_add_ops) cudaFree(t.d_add_ops);
            }
            delete g_extra[i].second;
            g_extra.erase(g_extra.begin() + i);
            return;
        }
}
} // namespace

static void side_geom(const LevelState& lv, const PatchState& ps, int axis, int* alo, int* ahi, int* ilo, int* ihi)
{
    for (int d = 0; d < 3; ++d)
    {
        if (d < lv.ndim)
        {
            ilo[d] = ps.lower[d];
            ihi[d] = ps.upper[d] + (d == axis ? 1 : 0);
            alo[d] = ilo[d] - lv.gcw[d];
            ahi[d] = ihi[d] + lv.gcw[d];
        }
        else
        {
            alo[d] = ahi[d] = ilo[d] = ihi[d] = 0;
        }
    }
}

namespace
{
struct HBox
{
    int lo[3], hi[3];
};
bool hbox_intersect(const HBox& a, const HBox& b, HBox& r)
{
    for (int d = 0; d < 3; ++d)
    {
        r.lo[d] = std::max(a.lo[d], b.lo[d]);
        r.hi[d] = std::min(a.hi[d], b.hi[d]);
        if (r.hi[d] < r.lo[d]) return false;
    }
    return true;
}
// a \ b as disjoint boxes (highest dimension first)
void hbox_minus(const HBox& a, const HBox& b, std::vector<HBox>& out)
{
    HBox in;
    if (!hbox_intersect(a, b, in))
    {
        out.push_back(a);
        return;
    }
    HBox cur = a;
    for (int d = 2; d >= 0; --d)
    {
        if (cur.lo[d] < in.lo[d])
        {
            HBox p = cur;
            p.hi[d] = in.lo[d] - 1;
            out.push_back(p);
        }
        if (cur.hi[d] > in.hi[d])
        {
            HBox p = cur;
            p.lo[d] = in.hi[d] + 1;
            out.push_back(p);
        }
        cur.lo[d] = in.lo[d];
        cur.hi[d] = in.hi[d];
    }
}
} // namespace

static unsigned halo_blocks(const LevelState& lv, const PatchState& ps, int axis);
static int build_halo_plan(ibk_ctx* ctx)
{
    LevelState& lv = ctx->lv;
    LevelExtra* ex = extra_of(ctx, true);
    HaloPlan& hp = ex->halo;
    const int ndim = lv.ndim, P = (int)lv.patches.size();
    int N[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d) N[d] = lv.domain_upper[d] - lv.domain_lower[d] + 1;
    for (int which = 0; which < 2; ++which) hp.d_args[which].assign((size_t)P * ndim, nullptr);
    std::vector<RegionMulti> ops[3]; // fill, accumulate, zero (all patches and axes)
    std::vector<FacePair> all_pairs;
    std::vector<FaceLayerJob> face_jobs;
    hp.d_face_save.assign((size_t)P * ndim, nullptr);
    hp.d_pairs.assign(ndim, nullptr);
    hp.n_pairs.assign(ndim, 0);
    for (int axis = 0; axis < ndim; ++axis)
    {
        std::vector<FacePair> pairs;
        for (int p = 0; p < P; ++p)
        {
            const PatchState& ps = lv.patches[p];
            int alo[3], ahi[3], ilo[3], ihi[3];
            side_geom(lv, ps, axis, alo, ahi, ilo, ihi);
            for (int which = 0; which < 2; ++which)
            {
                HaloArgs A;
                std::memset(&A, 0, sizeof(A));
                A.dst = which == 0 ? ps.u[axis] : ps.f[axis];
                A.pitch = ps.pitch[axis];
                A.n1 = ps.n[axis][1];
                A.ndim = ndim;
                A.rim = 0;
                for (int d = 0; d < 3; ++d)
                {
                    A.alo[d] = alo[d];
                    A.ahi[d] = ahi[d];
                    A.ilo[d] = ilo[d];
                    A.ihi[d] = ihi[d];
                    if (d < ndim) A.rim = std::max(A.rim, lv.gcw[d]);
                }
                // canonical source order: patch number, then periodic offset (o2, o1, o0)
                for (int q = 0; q < P; ++q)
                {
                    const PatchState& qs = lv.patches[q];
                    int qalo[3], qahi[3], qilo[3], qihi[3];
                    side_geom(lv, qs, axis, qalo, qahi, qilo, qihi);
                    int omin[3] = { 0, 0, 0 }, omax[3] = { 0, 0, 0 };
                    for (int d = 0; d < ndim; ++d)
                        if (lv.periodic[d])
                        {
                            omin[d] = -1;
                            omax[d] = 1;
                        }
                    for (int o2 = omin[2]; o2 <= omax[2]; ++o2)
                        for (int o1 = omin[1]; o1 <= omax[1]; ++o1)
                            for (int o0 = omin[0]; o0 <= omax[0]; ++o0)
                            {
                                if (q == p && o0 == 0 && o1 == 0 && o2 == 0) continue;
                                const int o[3] = { o0 * N[0], o1 * N[1], o2 * N[2] };
                                // does the shifted source array intersect the destination array at all?
                                bool hit = true;
                                for (int d = 0; d < ndim; ++d)
                                    hit = hit && (qalo[d] + o[d] <= ahi[d]) && (qahi[d] + o[d] >= alo[d]);
                                if (!hit) continue;
                                if (A.nsrc >= HALO_MAXSRC) return fail(ctx, IBK_ERR_INVALID, "too many halo sources for one patch");
                                HaloSrc& S = A.src[A.nsrc++];
                                S.ptr = which == 0 ? qs.u[axis] : qs.f[axis];
                                S.pitch = qs.pitch[axis];
                                S.n1 = qs.n[axis][1];
                                for (int d = 0; d < 3; ++d)
                                {
                                    S.alo[d] = qalo[d] + (d < ndim ? o[d] : 0);
                                    S.ahi[d] = qahi[d] + (d < ndim ? o[d] : 0);
                                    S.ilo[d] = qilo[d] + (d < ndim ? o[d] : 0);
                                    S.ihi[d] = qihi[d] + (d < ndim ? o[d] : 0);
                                }
                                // shared faces (f only, found once per unordered pair): q's LOWER face == p's UPPER face
                                if (which == 1 && qilo[axis] + o[axis] == ihi[axis])
                                {
                                    FacePair fp;
                                    std::memset(&fp, 0, sizeof(fp));
                                    bool ok = true;
                                    int lo[3], hi[3];
                                    for (int d = 0; d < 3; ++d)
                                    {
                                        if (d >= ndim)
                                        {
                                            lo[d] = hi[d] = 0;
                                            continue;
                                        }
                                        if (d == axis)
                                        {
                                            lo[d] = hi[d] = ihi[axis];
                                        }
                                        else
                                        {
                                            lo[d] = std::max(ilo[d], qilo[d] + o[d]);
                                            hi[d] = std::min(ihi[d], qihi[d] + o[d]);
                                        }
                                        ok = ok && hi[d] >= lo[d];
                                    }
                                    if (ok)
                                    {
                                        fp.a = ps.f[axis];
                                        fp.b = qs.f[axis];
                                        fp.a_pitch = ps.pitch[axis];
                                        fp.b_pitch = qs.pitch[axis];
                                        fp.a_n1 = ps.n[axis][1];
                                        fp.b_n1 = qs.n[axis][1];
                                        for (int d = 0; d < 3; ++d)
                                        {
                                            fp.a_off[d] = lo[d] - alo[d];
                                            fp.b_off[d] = lo[d] - (qalo[d] + (d < ndim ? o[d] : 0));
                                            fp.ext[d] = hi[d] - lo[d] + 1;
                                        }
                                        pairs.push_back(fp);
                                    }
                                }
                            }
                }
                // ---- the same operations as box lists (region_items_kernel)
                {
                    HBox arr, inter;
                    for (int d = 0; d < 3; ++d)
                    {
                        arr.lo[d] = alo[d];
                        arr.hi[d] = ahi[d];
                        inter.lo[d] = ilo[d];
                        inter.hi[d] = ihi[d];
                    }
                    auto make_op = [&](const HBox& r) {
                        RegionMulti op;
                        std::memset(&op, 0, sizeof(op));
                        op.dst = A.dst;
                        op.dst_pitch = A.pitch;
                        op.dst_n1 = A.n1;
                        for (int d = 0; d < 3; ++d)
                        {
                            op.dst_off[d] = r.lo[d] - alo[d];
                            op.ext[d] = r.hi[d] - r.lo[d] + 1;
                        }
                        return op;
                    };
                    auto add_src = [&](RegionMulti& op, const HBox& r, const HaloSrc& S) {
                        const int c = op.nsrc++;
                        op.src[c] = S.ptr;
                        op.src_pitch[c] = S.pitch;
                        op.src_n1[c] = S.n1;
                        for (int d = 0; d < 3; ++d) op.src_off[c][d] = r.lo[d] - S.alo[d];
                    };
                    std::vector<HBox> ghost;
                    hbox_minus(arr, inter, ghost);
                    if (which == 0)
                    {
                        // fill: every ghost element from the FIRST source whose interior holds it: what a source covers is taken
                        // out of the remaining region
                        std::vector<HBox> remaining = ghost;
                        for (int sidx = 0; sidx < A.nsrc && !remaining.empty(); ++sidx)
                        {
                            const HaloSrc& S = A.src[sidx];
                            HBox sint;
                            for (int d = 0; d < 3; ++d)
                            {
                                sint.lo[d] = S.ilo[d];
                                sint.hi[d] = S.ihi[d];
                            }
                            std::vector<HBox> next;
                            for (const HBox& R : remaining)
                            {
                                HBox I;
                                if (!hbox_intersect(R, sint, I))
                                {
                                    next.push_back(R);
                                    continue;
                                }
                                RegionMulti op = make_op(I);
                                add_src(op, I, S);
                                ops[0].push_back(op);
                                hbox_minus(R, I, next);
                            }
                            remaining.swap(next);
                        }
                    }
                    else
                    {
                        // zero of the ghost regions
                        for (const HBox& R : ghost) ops[2].push_back(make_op(R));
                        // accumulation: the interior adds the GHOST copies every source holds of it (interior copies of shared faces
                        // are face_sync_kernel's).  The rim of the interior is cut into boxes that each have ONE ordered list of
                        // sources (canonical order), so that an element's additions happen in one thread, in that order.
                        std::vector<std::pair<HBox, std::vector<int>>> cells;
                        for (int sidx = 0; sidx < A.nsrc; ++sidx)
                        {
                            const HaloSrc& S = A.src[sidx];
                            HBox sarr, sint, I;
                            for (int d = 0; d < 3; ++d)
                            {
                                sarr.lo[d] = S.alo[d];
                                sarr.hi[d] = S.ahi[d];
                                sint.lo[d] = S.ilo[d];
                                sint.hi[d] = S.ihi[d];
                            }
                            if (!hbox_intersect(inter, sarr, I)) continue;
                            std::vector<HBox> parts;
                            hbox_minus(I, sint, parts);
                            for (const HBox& B : parts)
                            {
                                std::vector<std::pair<HBox, std::vector<int>>> next;
                                std::vector<HBox> uncovered{ B };
                                for (auto& cell : cells)
                                {
                                    HBox J;
                                    if (!hbox_intersect(cell.first, B, J))
                                    {
                                        next.push_back(cell);
                                        continue;
                                    }
                                    std::vector<int> lst = cell.second;
                                    lst.push_back(sidx);
                                    next.push_back({ J, lst });
                                    std::vector<HBox> rest;
                                    hbox_minus(cell.first, J, rest);
                                    for (const HBox& R : rest) next.push_back({ R, cell.second });
                                    std::vector<HBox> unc2;
                                    for (const HBox& U : uncovered) hbox_minus(U, J, unc2);
                                    uncovered.swap(unc2);
                                }
                                for (const HBox& U : uncovered) next.push_back({ U, std::vector<int>{ sidx } });
                                cells.swap(next);
                            }
                        }
                        for (auto& cell : cells)
                        {
                            if ((int)cell.second.size() > REGION_MAXSRC) return fail(ctx, IBK_ERR_INVALID, "too many ghost copies of one region");
                            RegionMulti op = make_op(cell.first);
                            for (int sidx : cell.second) add_src(op, cell.first, A.src[sidx]);
                            ops[1].push_back(op);
                        }
                    }
                }
                HaloArgs* d_A = nullptr;
                CK(cudaMalloc(&d_A, sizeof(HaloArgs)));
                hp.allocs.push_back(d_A);
                if (which == 1)
                {
                    long long per = 1;
                    for (int d = 0; d < ndim; ++d) per *= (d == axis) ? 1 : (ihi[d] - ilo[d] + 1);
                    double* d_save = nullptr;
                    CK(cudaMalloc(&d_save, sizeof(double) * 2 * (size_t)per));
                    hp.allocs.push_back(d_save);
                    hp.d_face_save[(size_t)p * ndim + axis] = d_save;
                }
                CK(cudaMemcpy(d_A, &A, sizeof(HaloArgs), cudaMemcpyHostToDevice));
                hp.d_args[which][(size_t)p * ndim + axis] = d_A;
                if (which == 1)
                {
                    FaceLayerJob job;
                    job.args = d_A;
                    job.axis = axis;
                    job.save = hp.d_face_save[(size_t)p * ndim + axis];
                    face_jobs.push_back(job);
                    hp.face_job_blocks = std::max(hp.face_job_blocks, halo_blocks(lv, ps, axis));
                }
            }
        }
        all_pairs.insert(all_pairs.end(), pairs.begin(), pairs.end());
        hp.n_pairs[axis] = (int)pairs.size();
        if (!pairs.empty())
        {
            FacePair* d_p = nullptr;
            CK(cudaMalloc(&d_p, sizeof(FacePair) * pairs.size()));
            hp.allocs.push_back(d_p);
            CK(cudaMemcpy(d_p, pairs.data(), sizeof(FacePair) * pairs.size(), cudaMemcpyHostToDevice));
            hp.d_pairs[axis] = d_p;
        }
    }
    // CTAs in proportion to the size of the items: 16 warp work items (one row chunk each) per CTA
    for (int t = 0; t < 3; ++t)
    {
        std::vector<RegionMulti>& v = ops[t];
        unsigned nb = 0;
        for (RegionMulti& op : v)
        {
            op.block0 = nb;
            op.nblocks = region_blocks(op.ext);
            nb += op.nblocks;
        }
        hp.n_ops[t] = (int)v.size();
        hp.n_blocks[t] = nb;
        if (!v.empty())
        {
            CK(cudaMalloc(&hp.d_ops[t], sizeof(RegionMulti) * v.size()));
            hp.allocs.push_back(hp.d_ops[t]);
            CK(cudaMemcpy(hp.d_ops[t], v.data(), sizeof(RegionMulti) * v.size(), cudaMemcpyHostToDevice));
        }
    }
    hp.n_all_pairs = (int)all_pairs.size();
    if (!all_pairs.empty())
    {
        CK(cudaMalloc(&hp.d_all_pairs, sizeof(FacePair) * all_pairs.size()));
        hp.allocs.push_back(hp.d_all_pairs);
        CK(cudaMemcpy(hp.d_all_pairs, all_pairs.data(), sizeof(FacePair) * all_pairs.size(), cudaMemcpyHostToDevice));
    }
    hp.n_face_jobs = (int)face_jobs.size();
    if (!face_jobs.empty())
    {
        CK(cudaMalloc(&hp.d_face_jobs, sizeof(FaceLayerJob) * face_jobs.size()));
        hp.allocs.push_back(hp.d_face_jobs);
        CK(cudaMemcpy(hp.d_face_jobs, face_jobs.data(), sizeof(FaceLayerJob) * face_jobs.size(), cudaMemcpyHostToDevice));
    }
    return IBK_OK;
}

// all box operations of one kind in one launch
template <int MODE>
static cudaError_t launch_region_items(ibk_ctx* ctx, const HaloPlan& hp, int t)
{
    if (hp.n_ops[t] <= 0) return cudaSuccess;
    region_items_kernel<MODE><<<hp.n_blocks[t], 256, 0, ctx->L.stream>>>(hp.d_ops[t], hp.n_ops[t]);
    ctx->L.launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// level
// ---------------------------------------------------------------------------------------------
extern "C" int ibk_level_destroy(ibk_ctx* ctx)
{
    if (!ctx) return IBK_ERR_INVALID;
    LevelState& lv = ctx->lv;
    if (!lv.valid) return IBK_OK;
    cudaStreamSynchronize(ctx->L.stream);
    if (ctx->xfer_created)
    {
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->s_out);
        for (int w = 0; w < 2; ++w) ctx->pend_in[w] = ctx->pend_out[w] = false;
    }
    for (auto& ps : lv.patches)
        for (int a = 0; a < 3; ++a)
        {
            if (ps.u[a]) cudaFree(ps.u[a]);
            if (ps.f[a]) cudaFree(ps.f[a]);
        }
    if (lv.d_bins) cudaFree(lv.d_bins);
    for (void* p : { (void*)lv.X, (void*)lv.U, (void*)lv.F, (void*)lv.tmp, (void*)lv.lag, (void*)lv.lag_prev,
                     (void*)lv.gid, (void*)lv.cells, (void*)lv.owner, (void*)lv.escaped })
        if (p) cudaFree(p);
    for (double* p : lv.extra)
        if (p) cudaFree(p);
    for (double* p : lv.shadow)
        if (p) cudaFree(p);
    for (void* p : lv.force_allocs) cudaFree(p);
    if (lv.pos_of_id) cudaFree(lv.pos_of_id);
    if (lv.d_missing) cudaFree(lv.d_missing);
    bins_free(lv.bins);
    extra_drop(ctx);
    lv = LevelState();
    return IBK_OK;
}

extern "C" int ibk_level_create(ibk_ctx* ctx, const ibk_level_desc* desc)
{
    if (!ctx || !desc) return IBK_ERR_INVALID;
    if (desc->ndim != 2 && desc->ndim != 3) return fail(ctx, IBK_ERR_INVALID, "ndim must be 2 or 3");
    if (desc->n_patches < 0 || (desc->n_patches > 0 && (!desc->patch_lower || !desc->patch_upper)))
        return fail(ctx, IBK_ERR_INVALID, "patch boxes missing");
    ibk_level_destroy(ctx);
    LevelState& lv = ctx->lv;
    const int ndim = desc->ndim;
    lv.ndim = ndim;
    int gmax = 0;
    for (int d = 0; d < 3; ++d)
    {
        lv.domain_lower[d] = d < ndim ? desc->domain_lower[d] : 0;
        lv.domain_upper[d] = d < ndim ? desc->domain_upper[d] : 0;
        lv.periodic[d] = d < ndim ? desc->periodic[d] : 0;
        lv.gcw[d] = d < ndim ? desc->gcw[d] : 0;
        lv.x_lower[d] = d < ndim ? desc->x_lower[d] : 0.0;
        lv.x_upper[d] = d < ndim ? desc->x_upper[d] : 1.0;
        const int ncell = lv.domain_upper[d] - lv.domain_lower[d] + 1;
        // level dx as IndexUtilities::getCellIndex(X, grid_geom, ratio) forms it (dx0/ratio);
        // the caller passes the level's own domain box so dx = L / n_cells
        lv.dx[d] = d < ndim ? (lv.x_upper[d] - lv.x_lower[d]) / (double)ncell : 1.0;
        gmax = std::max(gmax, lv.gcw[d]);
    }
    lv.G = gmax + 4;
    int brick_base = 0;
    for (int p = 0; p < desc->n_patches; ++p)
    {
        PatchState ps;
        std::memset(&ps, 0, sizeof(ps));
        for (int d = 0; d < 3; ++d)
        {
            ps.lower[d] = d < ndim ? desc->patch_lower[p * ndim + d] : 0;
            ps.upper[d] = d < ndim ? desc->patch_upper[p * ndim + d] : 0;
        }
        fill_patch_bin(ps.pb, ndim, ps.lower, ps.upper, ps.lower, ps.upper, lv.G, brick_base);
        brick_base += ps.pb.nbricks;
        for (int a = 0; a < ndim; ++a)
        {
            for (int d = 0; d < 3; ++d)
                ps.n[a][d] = d < ndim ? ps.upper[d] - ps.lower[d] + 1 + 2 * lv.gcw[d] + (d == a ? 1 : 0) : 1;
            ps.pitch[a] = round_pitch(ps.n[a][0]);
            ps.elems[a] = (size_t)ps.pitch[a] * ps.n[a][1] * ps.n[a][2];
            CK(cudaMalloc(&ps.u[a], sizeof(double) * ps.elems[a]));
            CK(cudaMalloc(&ps.f[a], sizeof(double) * ps.elems[a]));
            CK(cudaMemsetAsync(ps.u[a], 0, sizeof(double) * ps.elems[a], ctx->L.stream));
            CK(cudaMemsetAsync(ps.f[a], 0, sizeof(double) * ps.elems[a], ctx->L.stream));
        }
        lv.patches.push_back(ps);
        lv.h_bins.push_back(ps.pb);
    }
    if (!lv.h_bins.empty())
    {
        CK(cudaMalloc(&lv.d_bins, sizeof(PatchBin) * lv.h_bins.size()));
        CK(cudaMemcpy(lv.d_bins, lv.h_bins.data(), sizeof(PatchBin) * lv.h_bins.size(), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&lv.escaped, sizeof(int)));
    lv.valid = true;
    // tile parameters per patch
    LevelExtra* ex = extra_of(ctx, true);
    ex->tp.resize(2 * lv.patches.size());
    for (size_t p = 0; p < lv.patches.size(); ++p)
    {
        const PatchState& ps = lv.patches[p];
        double xl[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
        int nvar[3] = { 1, 1, 1 };
        for (int d = 0; d < ndim; ++d)
        {
            // patch x_lower as SAMRAI's CartesianPatchGeometry holds it: x_lo + dx * (lower - domain_lower)
            const double pxl = lv.x_lower[d] + lv.dx[d] * (double)(ps.lower[d] - lv.domain_lower[d]);
            xl[d][0] = pxl;
            xl[d][1] = pxl - 0.5 * lv.dx[d];
            nvar[d] = 2;
        }
        for (int which = 0; which < 2; ++which)
        {
            ArrayComp comps[3];
            for (int a = 0; a < ndim; ++a)
            {
                comps[a].ptr = which == 0 ? ps.u[a] : ps.f[a];
                comps[a].pitch = ps.pitch[a];
                comps[a].vcol = a;
                comps[a].axis = a;
                for (int d = 0; d < 3; ++d)
                {
                    comps[a].n[d] = ps.n[a][d];
                    comps[a].nugc[d] = d < ndim ? lv.gcw[d] : 0;
                    comps[a].var[d] = (d == a) ? 1 : 0;
                }
            }
            make_tile_params(ex->tp[2 * p + which], ndim, lv.dx, xl, nvar, ps.pb, ndim, comps);
        }
    }
    int rc = build_halo_plan(ctx);
    if (rc != IBK_OK) return rc;
    return IBK_OK;
}

static int check_patch_axis(ibk_ctx* ctx, int which, int patch, int axis)
{
    if (which < 0 || which > 1 || patch < 0 || patch >= (int)ctx->lv.patches.size() || axis < 0 || axis >= ctx->lv.ndim)
        return fail(ctx, IBK_ERR_INVALID, "bad (which, patch, axis)");
    return IBK_OK;
}

constexpr size_t STAGE_BYTES = 64u << 20; // dense row blocks of the grid transfers (see copy_dense_to_pitched)

// ---- asynchronous transfers: uploads and downloads of grid data on their own streams, so that the PCIe
// traffic of one array overlaps the kernels (and the opposite-direction traffic) of another.
static cudaError_t xfer_init(ibk_ctx* ctx)
{
    if (ctx->xfer_created) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int w = 0; w < 2; ++w)
    {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_in[w], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ctx->ev_out[w], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->xfer_created = true;
    return cudaSuccess;
}
// `stream` waits for the asynchronous transfers in flight that touch array `which` (0 = u, 1 = f)
static cudaError_t grid_deps(ibk_ctx* ctx, int which, cudaStream_t stream)
{
    cudaError_t e;
    if (ctx->pend_in[which] && stream != ctx->s_in)
        if ((e = cudaStreamWaitEvent(stream, ctx->ev_in[which], 0)) != cudaSuccess) return e;
    if (ctx->pend_out[which] && stream != ctx->s_out)
        if ((e = cudaStreamWaitEvent(stream, ctx->ev_out[which], 0)) != cudaSuccess) return e;
    return cudaSuccess;
}
#define GRID_DEPS(which) CK(grid_deps(ctx, (which), ctx->L.stream))

extern "C" int ibk_grid_upload(ibk_ctx* ctx, int which, int patch, int axis, const double* h_data)
{
    NEED_LEVEL();
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    GRID_DEPS(which);
    PatchState& ps = ctx->lv.patches[patch];
    CK(ctx->b_stage[0].reserve(STAGE_BYTES));
    CK(copy_dense_to_pitched(ctx->L, h_data, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], ps.n[axis], ctx->lv.ndim,
                             cudaMemcpyHostToDevice, ctx->b_stage[0].p, STAGE_BYTES));
    return IBK_OK;
}
extern "C" int ibk_grid_download(ibk_ctx* ctx, int which, int patch, int axis, double* h_data)
{
    NEED_LEVEL();
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    GRID_DEPS(which);
    PatchState& ps = ctx->lv.patches[patch];
    CK(ctx->b_stage[0].reserve(STAGE_BYTES));
    CK(copy_pitched_to_dense(ctx->L, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], h_data, ps.n[axis], ctx->lv.ndim,
                             cudaMemcpyDeviceToHost, ctx->b_stage[0].p, STAGE_BYTES));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return IBK_OK;
}
extern "C" int ibk_grid_upload_async(ibk_ctx* ctx, int which, int patch, int axis, const double* h_data)
{
    NEED_LEVEL();
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    if (!h_data) return fail(ctx, IBK_ERR_INVALID, "null host pointer");
    CK(xfer_init(ctx));
    // the copy is ordered after the work already queued on the compute stream (which may still read the
    // array) and after a download of the same array that is in flight
    CK(cudaEventRecord(ctx->ev_order, ctx->L.stream));
    CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_order, 0));
    CK(grid_deps(ctx, which, ctx->s_in));
    PatchState& ps = ctx->lv.patches[patch];
    Launcher Lin = ctx->L;
    Lin.stream = ctx->s_in;
    CK(ctx->b_stage[1].reserve(STAGE_BYTES));
    CK(copy_dense_to_pitched(Lin, h_data, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], ps.n[axis], ctx->lv.ndim,
                             cudaMemcpyHostToDevice, ctx->b_stage[1].p, STAGE_BYTES));
    ctx->L.launches += Lin.launches - ctx->L.launches;
    CK(cudaEventRecord(ctx->ev_in[which], ctx->s_in));
    ctx->pend_in[which] = true;
    return IBK_OK;
}
extern "C" int ibk_grid_download_async(ibk_ctx* ctx, int which, int patch, int axis, double* h_data)
{
    NEED_LEVEL();
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    if (!h_data) return fail(ctx, IBK_ERR_INVALID, "null host pointer");
    CK(xfer_init(ctx));
    CK(cudaEventRecord(ctx->ev_order, ctx->L.stream));
    CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_order, 0));
    CK(grid_deps(ctx, which, ctx->s_out));
    PatchState& ps = ctx->lv.patches[patch];
    Launcher Lout = ctx->L;
    Lout.stream = ctx->s_out;
    CK(ctx->b_stage[2].reserve(STAGE_BYTES));
    CK(copy_pitched_to_dense(Lout, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], h_data, ps.n[axis], ctx->lv.ndim,
                             cudaMemcpyDeviceToHost, ctx->b_stage[2].p, STAGE_BYTES));
    ctx->L.launches += Lout.launches - ctx->L.launches;
    CK(cudaEventRecord(ctx->ev_out[which], ctx->s_out));
    ctx->pend_out[which] = true;
    return IBK_OK;
}
extern "C" int ibk_transfers_wait(ibk_ctx* ctx)
{
    if (!ctx) return IBK_ERR_INVALID;
    if (ctx->xfer_created)
    {
        CK(cudaStreamSynchronize(ctx->s_in));
        CK(cudaStreamSynchronize(ctx->s_out));
    }
    for (int w = 0; w < 2; ++w) ctx->pend_in[w] = ctx->pend_out[w] = false;
    return IBK_OK;
}
extern "C" int ibk_grid_fill(ibk_ctx* ctx, int which, double value)
{
    NEED_LEVEL();
    if (which < 0 || which > 1) return fail(ctx, IBK_ERR_INVALID, "which must be 0 (u) or 1 (f)");
    GRID_DEPS(which);
    for (auto& ps : ctx->lv.patches)
        for (int a = 0; a < ctx->lv.ndim; ++a) CK(launch_fill(ctx->L, which == 0 ? ps.u[a] : ps.f[a], ps.elems[a], value));
    return IBK_OK;
}
extern "C" int ibk_grid_device_ptr(ibk_ctx* ctx, int which, int patch, int axis, double** d_ptr, long long* pitch, int* dims)
{
    NEED_LEVEL();
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    GRID_DEPS(which);
    PatchState& ps = ctx->lv.patches[patch];
    if (d_ptr) *d_ptr = which == 0 ? ps.u[axis] : ps.f[axis];
    if (pitch) *pitch = ps.pitch[axis];
    if (dims)
        for (int d = 0; d < ctx->lv.ndim; ++d) dims[d] = ps.n[axis][d];
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// markers
// ---------------------------------------------------------------------------------------------
// Capacity for n markers; the first n_keep entries of every column survive a reallocation.
static int reserve_markers(ibk_ctx* ctx, int n, int n_keep = 0)
{
    LevelState& lv = ctx->lv;
    if ((long long)n <= lv.stride) return IBK_OK;
    const long long stride = ((long long)std::max(n, 1024) + 31) / 32 * 32;
    const int ndim = lv.ndim;
    for (double** pp : { &lv.X, &lv.U, &lv.F, &lv.tmp, &lv.extra[0], &lv.extra[1], &lv.extra[2] })
    {
        if (!*pp && (pp == &lv.extra[0] || pp == &lv.extra[1] || pp == &lv.extra[2])) continue; // allocated on first use
        double* fresh = nullptr;
        CK(cudaMalloc(&fresh, sizeof(double) * (size_t)stride * ndim));
        CK(cudaMemsetAsync(fresh, 0, sizeof(double) * (size_t)stride * ndim, ctx->L.stream));
        if (*pp && n_keep > 0 && pp != &lv.tmp)
            CK(cudaMemcpy2DAsync(fresh, sizeof(double) * (size_t)stride, *pp, sizeof(double) * (size_t)lv.stride,
                                 sizeof(double) * (size_t)n_keep, ndim, cudaMemcpyDeviceToDevice, ctx->L.stream));
        if (*pp)
        {
            CK(cudaStreamSynchronize(ctx->L.stream));
            cudaFree(*pp);
        }
        *pp = fresh;
    }
    for (uint32_t** pp : { &lv.lag, &lv.lag_prev, &lv.gid })
    {
        if (pp == &lv.gid && !lv.gid) continue; // global ids only once they were set
        uint32_t* fresh = nullptr;
        CK(cudaMalloc(&fresh, sizeof(uint32_t) * (size_t)stride));
        if (*pp && n_keep > 0)
            CK(cudaMemcpyAsync(fresh, *pp, sizeof(uint32_t) * (size_t)n_keep, cudaMemcpyDeviceToDevice, ctx->L.stream));
        if (*pp)
        {
            CK(cudaStreamSynchronize(ctx->L.stream));
            cudaFree(*pp);
        }
        *pp = fresh;
    }
    if (lv.cells) cudaFree(lv.cells);
    if (lv.owner) cudaFree(lv.owner);
    CK(cudaMalloc(&lv.cells, sizeof(int) * (size_t)stride * ndim));
    CK(cudaMalloc(&lv.owner, sizeof(int) * (size_t)stride));
    lv.stride = stride;
    lv.pos_valid = false;
    return IBK_OK;
}

constexpr int IBK_NCOLS = 6; // X, U, F, X_current, X_new, auxiliary
static double* column_of(LevelState& lv, int which)
{
    return which == 0 ? lv.X : which == 1 ? lv.U : which == 2 ? lv.F : lv.extra[which - 3];
}
// columns 3..5 exist once they are used
static int ensure_column(ibk_ctx* ctx, int which)
{
    LevelState& lv = ctx->lv;
    if (which < 0 || which >= IBK_NCOLS) return fail(ctx, IBK_ERR_INVALID, "bad marker column");
    if (which < 3 || lv.extra[which - 3] || lv.stride == 0) return IBK_OK;
    CK(cudaMalloc(&lv.extra[which - 3], sizeof(double) * (size_t)lv.stride * lv.ndim));
    CK(cudaMemsetAsync(lv.extra[which - 3], 0, sizeof(double) * (size_t)lv.stride * lv.ndim, ctx->L.stream));
    return IBK_OK;
}

extern "C" int ibk_markers_upload(ibk_ctx* ctx, int which, const double* h_data)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (which < 0 || which >= IBK_NCOLS || !h_data) return fail(ctx, IBK_ERR_INVALID, "bad marker column");
    if (lv.n <= 0) return IBK_OK;
    if (int rc = ensure_column(ctx, which)) return rc;
    const size_t bytes = sizeof(double) * (size_t)lv.n * lv.ndim;
    CK(ctx->b_io[5].reserve(bytes));
    CK(cudaMemcpyAsync(ctx->b_io[5].p, h_data, bytes, cudaMemcpyHostToDevice, ctx->L.stream));
    // AoS (Lagrangian order) -> SoA (Lagrangian order) -> storage order: col[i] = lagcol[lag[i]]
    CK(aos_to_soa(ctx->L, ctx->b_io[5].as<double>(), lv.tmp, lv.stride, lv.n, lv.ndim));
    CK(gather_columns(ctx->L, lv.tmp, lv.stride, column_of(lv, which), lv.stride, lv.lag, lv.n, lv.ndim));
    if (which == 0) lv.binned = false;
    return IBK_OK;
}
extern "C" int ibk_markers_download(ibk_ctx* ctx, int which, double* h_data)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (which < 0 || which >= IBK_NCOLS || !h_data) return fail(ctx, IBK_ERR_INVALID, "bad marker column");
    if (lv.n <= 0) return IBK_OK;
    if (int rc = ensure_column(ctx, which)) return rc;
    const size_t bytes = sizeof(double) * (size_t)lv.n * lv.ndim;
    CK(ctx->b_io[5].reserve(bytes));
    CK(scatter_columns(ctx->L, column_of(lv, which), lv.stride, lv.tmp, lv.stride, lv.lag, lv.n, lv.ndim));
    CK(soa_to_aos(ctx->L, lv.tmp, lv.stride, ctx->b_io[5].as<double>(), lv.n, lv.ndim));
    CK(cudaMemcpyAsync(h_data, ctx->b_io[5].p, bytes, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return IBK_OK;
}
extern "C" int ibk_markers_set_positions(ibk_ctx* ctx, const double* h_X, int n_markers)
{
    NEED_LEVEL();
    if (n_markers < 0 || (n_markers > 0 && !h_X)) return fail(ctx, IBK_ERR_INVALID, "bad marker positions");
    LevelState& lv = ctx->lv;
    if (int rc = reserve_markers(ctx, n_markers)) return rc;
    lv.n = n_markers;
    lv.binned = false;
    if (lv.gid) cudaFree(lv.gid); // the numbering is reset to 0..n-1 (host rows = Lagrangian indices)
    lv.gid = nullptr;
    lv.id_bound = 0;
    lv.mig_n = -1;
    lv.pos_valid = false;
    if (n_markers == 0) return IBK_OK;
    iota_kernel<<<(n_markers + 255) / 256, 256, 0, ctx->L.stream>>>(lv.lag, n_markers);
    ctx->L.launches++;
    return ibk_markers_upload(ctx, 0, h_X);
}
extern "C" int ibk_markers_count(const ibk_ctx* ctx)
{
    return (ctx && ctx->lv.valid) ? ctx->lv.n : 0;
}
extern "C" int ibk_markers_device_ptr(ibk_ctx* ctx, int which, double** d_ptr, long long* stride)
{
    NEED_LEVEL();
    if (which < 0 || which >= IBK_NCOLS) return fail(ctx, IBK_ERR_INVALID, "bad marker column");
    if (int rc = ensure_column(ctx, which)) return rc;
    if (d_ptr) *d_ptr = column_of(ctx->lv, which);
    if (stride) *stride = ctx->lv.stride;
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// rebin
// ---------------------------------------------------------------------------------------------
extern "C" int ibk_rebin(ibk_ctx* ctx, int error_if_points_leave_domain)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    const int ndim = lv.ndim, n = lv.n;
    if (ctx->timing) CK(cudaEventRecord(ctx->ev[2][0], ctx->L.stream));
    DomainGeom dg;
    std::memset(&dg, 0, sizeof(dg));
    dg.ndim = ndim;
    CellGeom cg;
    std::memset(&cg, 0, sizeof(cg));
    cg.ndim = ndim;
    cg.two_branch = 1; // IndexUtilities::getCellIndex(X, grid_geom, ratio), IndexUtilities-inl.h:225-242
    for (int d = 0; d < ndim; ++d)
    {
        dg.x_lower[d] = cg.x_lower[d] = lv.x_lower[d];
        dg.x_upper[d] = cg.x_upper[d] = lv.x_upper[d];
        dg.periodic[d] = lv.periodic[d];
        cg.dx[d] = lv.dx[d];
        cg.ilower[d] = lv.domain_lower[d];
        cg.iupper[d] = lv.domain_upper[d];
    }
    CK(cudaMemsetAsync(lv.escaped, 0, sizeof(int), ctx->L.stream));
    if (error_if_points_leave_domain)
    {
        // the reference aborts before it moves anything (LDataManager.cpp:1410-1416): check first, X untouched on failure
        int esc = 0;
        CK(wrap_positions(ctx->L, dg, lv.X, lv.stride, n, lv.escaped, /*check_only*/ true));
        CK(cudaMemcpyAsync(&esc, lv.escaped, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
        if (esc > 0) return fail(ctx, IBK_ERR_ESCAPED, "IB point has escaped from the computational domain!");
    }
    CK(wrap_positions(ctx->L, dg, lv.X, lv.stride, n, lv.escaped));
    if (n > 0) CK(cudaMemcpyAsync(lv.lag_prev, lv.lag, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->L.stream));
    // in-cell order = ascending Lagrangian index (LDataManager.cpp:1505): the global one when it is known
    CK(bins_build(lv.bins, ctx->L, cg, lv.d_bins, (int)lv.h_bins.size(), lv.h_bins.data(), lv.X, lv.stride,
                  lv.gid ? lv.gid : lv.lag, lv.gid ? std::max(lv.id_bound, 1u) : (uint32_t)std::max(n, 1), n, lv.cells,
                  lv.owner));
    lv.mig_n = -1;
    if (n > 0)
    {
        const uint32_t* perm = lv.bins.vals[lv.bins.sorted_in];
        // all columns in one launch, each into its shadow; then column and shadow swap roles (no copy back)
        if (lv.shadow_stride != lv.stride)
        {
            for (double*& sh : lv.shadow)
            {
                if (sh) cudaFree(sh);
                sh = nullptr;
            }
            lv.shadow_stride = lv.stride;
        }
        GatherSets gs;
        std::memset(&gs, 0, sizeof(gs));
        double** cols[6] = { &lv.X, &lv.U, &lv.F, &lv.extra[0], &lv.extra[1], &lv.extra[2] };
        int which_col[6];
        for (int k = 0; k < 6; ++k)
        {
            if (!*cols[k]) continue;
            if (!lv.shadow[k]) CK(cudaMalloc(&lv.shadow[k], sizeof(double) * (size_t)lv.stride * ndim));
            gs.in[gs.nsets] = *cols[k];
            gs.out[gs.nsets] = lv.shadow[k];
            which_col[gs.nsets++] = k;
        }
        CK(gather_column_sets(ctx->L, gs, lv.stride, perm, n, ndim));
        for (int q = 0; q < gs.nsets; ++q) std::swap(*cols[which_col[q]], lv.shadow[which_col[q]]);
        if (lv.gid)
        {
            CK(extract_low32(ctx->L, lv.bins.keys[lv.bins.sorted_in], lv.gid, n, lv.bins.tie_bits));
            CK(gather_u32(ctx->L, lv.lag_prev, perm, n, lv.lag));
        }
        else
            CK(extract_low32(ctx->L, lv.bins.keys[lv.bins.sorted_in], lv.lag, n, lv.bins.tie_bits));
    }
    lv.binned = true;
    lv.pos_valid = false;
    if (ctx->timing)
    {
        CK(cudaEventRecord(ctx->ev[2][1], ctx->L.stream));
        ctx->ev_valid[2] = true;
    }
    return IBK_OK;
}

// Markers of the context that a local patch accepted at the last ibk_rebin.  The others (their cell lies in no local
// patch: they belong to another rank, or the patches do not cover them) sit behind the binned ones and are skipped by
// spread, interpolation and the force kernels until ibk_migrate hands them over.
extern "C" int ibk_markers_owned_count(ibk_ctx* ctx, int* n_owned)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!n_owned) return fail(ctx, IBK_ERR_INVALID, "null pointer");
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "ibk_rebin has not run");
    *n_owned = 0;
    if (lv.n == 0 || !lv.bins.brick_start) return IBK_OK;
    CK(cudaMemcpyAsync(n_owned, lv.bins.brick_start + lv.bins.total_bricks, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return IBK_OK;
}

extern "C" int ibk_bin_get_cells(ibk_ctx* ctx, int* h_cells, int* h_owner)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "ibk_rebin has not run");
    const int n = lv.n, ndim = lv.ndim;
    if (n == 0) return IBK_OK;
    std::vector<int> cells((size_t)n * ndim), owner(n);
    std::vector<uint32_t> lagp(n);
    CK(cudaMemcpyAsync(cells.data(), lv.cells, sizeof(int) * cells.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(owner.data(), lv.owner, sizeof(int) * owner.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(lagp.data(), lv.lag_prev, sizeof(uint32_t) * lagp.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int i = 0; i < n; ++i)
    {
        const uint32_t l = lagp[i];
        if (h_cells)
            for (int d = 0; d < ndim; ++d) h_cells[(size_t)l * ndim + d] = cells[(size_t)i * ndim + d];
        if (h_owner) h_owner[l] = owner[i];
    }
    return IBK_OK;
}
extern "C" int ibk_bin_get_order(ibk_ctx* ctx, int* h_lag_idx)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "ibk_rebin has not run");
    if (lv.n == 0 || !h_lag_idx) return IBK_OK;
    CK(cudaMemcpyAsync(h_lag_idx, lv.gid ? lv.gid : lv.lag, sizeof(uint32_t) * (size_t)lv.n, cudaMemcpyDeviceToHost,
                       ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// marker migration between ranks (LDataManager::endDataRedistribution's scatter, LDataManager.cpp:1824-1837)
// ---------------------------------------------------------------------------------------------
static int n_active_of(const LevelState& lv)
{
    int n_active = 0;
    for (int v : lv.bins.range_last) n_active = std::max(n_active, v);
    return n_active;
}
static int ceil_log2_u(uint32_t v)
{
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)v) ++b;
    return b;
}

extern "C" int ibk_markers_set_ids(ibk_ctx* ctx, const unsigned* h_ids, unsigned id_bound)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!h_ids && lv.n > 0) return fail(ctx, IBK_ERR_INVALID, "null id array");
    if (lv.n == 0)
    {
        lv.id_bound = id_bound;
        return IBK_OK;
    }
    if (!lv.gid) CK(cudaMalloc(&lv.gid, sizeof(uint32_t) * (size_t)lv.stride));
    CK(ctx->b_mig[1].reserve(sizeof(uint32_t) * (size_t)lv.n));
    CK(cudaMemcpyAsync(ctx->b_mig[1].p, h_ids, sizeof(uint32_t) * (size_t)lv.n, cudaMemcpyHostToDevice, ctx->L.stream));
    CK(gather_u32(ctx->L, ctx->b_mig[1].as<uint32_t>(), lv.lag, lv.n, lv.gid)); // host rows -> storage order
    CK(cudaStreamSynchronize(ctx->L.stream));
    lv.id_bound = id_bound;
    lv.binned = false;
    lv.pos_valid = false;
    return IBK_OK;
}
extern "C" int ibk_markers_get_ids(ibk_ctx* ctx, unsigned* h_ids)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (lv.n == 0) return IBK_OK;
    if (!h_ids) return fail(ctx, IBK_ERR_INVALID, "null id array");
    std::vector<uint32_t> row(lv.n), id(lv.n);
    CK(cudaMemcpyAsync(row.data(), lv.lag, sizeof(uint32_t) * (size_t)lv.n, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(id.data(), lv.gid ? lv.gid : lv.lag, sizeof(uint32_t) * (size_t)lv.n, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int i = 0; i < lv.n; ++i) h_ids[row[i]] = id[i];
    return IBK_OK;
}

extern "C" int ibk_migrate_plan(ibk_ctx* ctx, int n_patches, const int* patch_lower, const int* patch_upper, const int* patch_rank,
                                int n_ranks, int my_rank, int* h_send_counts)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "ibk_rebin has not run");
    if (n_patches <= 0 || !patch_lower || !patch_upper || !patch_rank || n_ranks <= 0 || my_rank < 0 || my_rank >= n_ranks ||
        !h_send_counts)
        return fail(ctx, IBK_ERR_INVALID, "bad migration arguments");
    const int ndim = lv.ndim;
    const int n_active = n_active_of(lv), n_tail = lv.n - n_active;
    for (int r = 0; r < n_ranks; ++r) h_send_counts[r] = 0;
    lv.mig_n = 0;
    lv.mig_order = nullptr;
    if (n_tail <= 0) return IBK_OK;
    if (!lv.gid) return fail(ctx, IBK_ERR_STATE, "markers leave this rank but no global indices were set (ibk_markers_set_ids)");
    if (lv.extra[0] || lv.extra[1] || lv.extra[2])
        return fail(ctx, IBK_ERR_STATE, "marker migration carries the columns X, U, F only (columns 3..5 are in use)");
    DevBuf* B = ctx->b_mig;
    CK(B[0].reserve(sizeof(uint64_t) * (size_t)n_tail));
    CK(B[1].reserve(sizeof(uint32_t) * (size_t)n_tail));
    CK(B[2].reserve(sizeof(uint64_t) * (size_t)n_tail));
    CK(B[3].reserve(sizeof(uint32_t) * (size_t)n_tail));
    CK(B[4].reserve(radix_sort_temp_bytes(n_tail)));
    CK(B[5].reserve(sizeof(int) * (size_t)n_patches * (2 * ndim + 1)));
    CK(B[6].reserve(sizeof(int) * (size_t)(n_ranks + 2)));
    int* d_plo = B[5].as<int>();
    int* d_phi = d_plo + (size_t)n_patches * ndim;
    int* d_prank = d_phi + (size_t)n_patches * ndim;
    CK(cudaMemcpyAsync(d_plo, patch_lower, sizeof(int) * (size_t)n_patches * ndim, cudaMemcpyHostToDevice, ctx->L.stream));
    CK(cudaMemcpyAsync(d_phi, patch_upper, sizeof(int) * (size_t)n_patches * ndim, cudaMemcpyHostToDevice, ctx->L.stream));
    CK(cudaMemcpyAsync(d_prank, patch_rank, sizeof(int) * (size_t)n_patches, cudaMemcpyHostToDevice, ctx->L.stream));
    CellGeom cg;
    std::memset(&cg, 0, sizeof(cg));
    cg.ndim = ndim;
    cg.two_branch = 1;
    for (int d = 0; d < ndim; ++d)
    {
        cg.x_lower[d] = lv.x_lower[d];
        cg.x_upper[d] = lv.x_upper[d];
        cg.dx[d] = lv.dx[d];
        cg.ilower[d] = lv.domain_lower[d];
        cg.iupper[d] = lv.domain_upper[d];
    }
    CK(migrate_dest(ctx->L, cg, d_plo, d_phi, d_prank, n_patches, n_ranks, lv.X, lv.stride, n_active, n_tail, B[0].as<uint64_t>(),
                    B[1].as<uint32_t>()));
    const int which = radix_sort_pairs(B[0].as<uint64_t>(), B[1].as<uint32_t>(), B[2].as<uint64_t>(), B[3].as<uint32_t>(), n_tail, 0,
                                       ceil_log2_u((uint32_t)n_ranks + 1), B[4].p, ctx->L.stream, &ctx->L.launches);
    const uint64_t* keys_sorted = which ? B[2].as<uint64_t>() : B[0].as<uint64_t>();
    lv.mig_order = which ? B[3].as<uint32_t>() : B[1].as<uint32_t>();
    CK(bucket_offsets(ctx->L, keys_sorted, n_tail, n_ranks + 1, B[6].as<int>()));
    std::vector<int> start(n_ranks + 2);
    CK(cudaMemcpyAsync(start.data(), B[6].p, sizeof(int) * start.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int r = 0; r < n_ranks; ++r) h_send_counts[r] = start[r + 1] - start[r];
    if (start[n_ranks + 1] - start[n_ranks] > 0)
        return fail(ctx, IBK_ERR_ESCAPED, "a marker lies in no patch of the level (markers must stay on the finest level)");
    if (h_send_counts[my_rank] > 0) return fail(ctx, IBK_ERR_STATE, "a marker of a local patch was not binned locally");
    lv.mig_n = start[n_ranks];
    return IBK_OK;
}
extern "C" int ibk_migrate_pack(ibk_ctx* ctx, double* d_buf)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (lv.mig_n < 0) return fail(ctx, IBK_ERR_STATE, "ibk_migrate_plan has not run");
    if (lv.mig_n == 0) return IBK_OK;
    if (!d_buf) return fail(ctx, IBK_ERR_INVALID, "null buffer");
    CK(migrate_pack(ctx->L, lv.mig_order, lv.mig_n, lv.X, lv.U, lv.F, lv.stride, lv.ndim, lv.gid, d_buf));
    return IBK_OK;
}
extern "C" int ibk_migrate_unpack(ibk_ctx* ctx, const double* d_buf, int n_recv, unsigned id_bound)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (lv.mig_n < 0) return fail(ctx, IBK_ERR_STATE, "ibk_migrate_plan has not run");
    if (n_recv < 0 || (n_recv > 0 && !d_buf)) return fail(ctx, IBK_ERR_INVALID, "bad receive buffer");
    const int n_active = n_active_of(lv);
    const int n_new = n_active + n_recv;
    lv.mig_n = -1;
    if (n_new == lv.n && n_recv == 0) return IBK_OK; // nothing left, nothing arrived
    if (n_recv > 0 && !lv.gid) return fail(ctx, IBK_ERR_STATE, "markers arrive but no global indices were set (ibk_markers_set_ids)");
    if (int rc = reserve_markers(ctx, n_new, n_active)) return rc;
    if (n_recv > 0) CK(migrate_append(ctx->L, d_buf, n_recv, n_active, lv.X, lv.U, lv.F, lv.stride, lv.ndim, lv.gid));
    lv.n = n_new;
    lv.id_bound = std::max(lv.id_bound, id_bound);
    lv.binned = false;
    lv.pos_valid = false;
    if (n_new > 0 && lv.gid)
    {
        // host rows := ascending global index among the markers now held (identity numbering for one rank)
        DevBuf* B = ctx->b_mig;
        CK(B[0].reserve(sizeof(uint64_t) * (size_t)n_new));
        CK(B[1].reserve(sizeof(uint32_t) * (size_t)n_new));
        CK(B[2].reserve(sizeof(uint64_t) * (size_t)n_new));
        CK(B[3].reserve(sizeof(uint32_t) * (size_t)n_new));
        CK(B[4].reserve(radix_sort_temp_bytes(n_new)));
        CK(id_keys(ctx->L, lv.gid, n_new, B[0].as<uint64_t>(), B[1].as<uint32_t>()));
        const int which = radix_sort_pairs(B[0].as<uint64_t>(), B[1].as<uint32_t>(), B[2].as<uint64_t>(), B[3].as<uint32_t>(), n_new,
                                           0, ceil_log2_u(std::max(lv.id_bound, 2u)), B[4].p, ctx->L.stream, &ctx->L.launches);
        CK(rank_scatter(ctx->L, which ? B[3].as<uint32_t>() : B[1].as<uint32_t>(), n_new, lv.lag));
    }
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// N1: Lagrangian forces and marker-column algebra on the device (IBStandardForceGen, IBMethod steps)
// ---------------------------------------------------------------------------------------------
template <class T>
static cudaError_t force_upload(LevelState& lv, const std::vector<T>& h, const T** d_out)
{
    *d_out = nullptr;
    if (h.empty()) return cudaSuccess;
    T* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(T) * h.size());
    if (e != cudaSuccess) return e;
    lv.force_allocs.push_back(d);
    *d_out = d;
    return cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
}
// node -> items CSR; items of one node ascend with the element number (the order the reference's loop meets them)
static void build_csr(int n_nodes, const std::vector<std::pair<int, int>>& node_item, std::vector<int>& ptr, std::vector<int>& items)
{
    ptr.assign((size_t)n_nodes + 1, 0);
    for (const auto& ni : node_item) ptr[(size_t)ni.first + 1]++;
    for (int l = 0; l < n_nodes; ++l) ptr[(size_t)l + 1] += ptr[l];
    items.resize(node_item.size());
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (const auto& ni : node_item) items[(size_t)fill[ni.first]++] = ni.second; // node_item is in element order
}
static int force_node_bound(const LevelState& lv)
{
    return lv.gid ? (int)lv.id_bound : lv.n;
}
static int check_nodes(ibk_ctx* ctx, int n, std::initializer_list<const int*> arrays, const char* what)
{
    const int bound = force_node_bound(ctx->lv);
    for (const int* a : arrays)
    {
        if (!a) return fail(ctx, IBK_ERR_INVALID, std::string("null index array for ") + what);
        for (int k = 0; k < n; ++k)
            if (a[k] < 0 || a[k] >= bound) return fail(ctx, IBK_ERR_INVALID, std::string(what) + ": node index out of range");
    }
    return IBK_OK;
}
// The CSR pointer arrays of the three groups must agree on n_nodes: rebuilt whenever a group is (re)set.
static int force_commit(ibk_ctx* ctx)
{
    ctx->lv.force.n_nodes = force_node_bound(ctx->lv);
    return IBK_OK;
}

extern "C" int ibk_force_clear(ibk_ctx* ctx)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (void* p : lv.force_allocs) cudaFree(p);
    lv.force_allocs.clear();
    lv.force = ForceTables();
    return IBK_OK;
}

extern "C" int ibk_force_set_springs(ibk_ctx* ctx, int n, const int* master, const int* slave, const double* kappa,
                                     const double* rest_length)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (n < 0 || (n > 0 && (!kappa || !rest_length))) return fail(ctx, IBK_ERR_INVALID, "bad spring arrays");
    if (int rc = check_nodes(ctx, n, { master, slave }, "springs")) return rc;
    for (int k = 0; k < n; ++k)
        if (master[k] == slave[k]) return fail(ctx, IBK_ERR_INVALID, "spring connects a node to itself"); // TBOX_ASSERT :852
    const int nn = force_node_bound(lv);
    std::vector<std::pair<int, int>> node_item;
    node_item.reserve(2 * (size_t)n);
    for (int k = 0; k < n; ++k)
    {
        node_item.push_back({ master[k], 2 * k });
        node_item.push_back({ slave[k], 2 * k + 1 });
    }
    std::vector<int> ptr, items;
    build_csr(nn, node_item, ptr, items);
    CK(force_upload(lv, ptr, &lv.force.spring_ptr));
    CK(force_upload(lv, items, &lv.force.spring_items));
    CK(force_upload(lv, std::vector<int>(master, master + n), &lv.force.spring_mastr));
    CK(force_upload(lv, std::vector<int>(slave, slave + n), &lv.force.spring_slave));
    CK(force_upload(lv, std::vector<double>(kappa, kappa + n), &lv.force.spring_kappa));
    CK(force_upload(lv, std::vector<double>(rest_length, rest_length + n), &lv.force.spring_rest));
    return force_commit(ctx);
}

extern "C" int ibk_force_set_beams(ibk_ctx* ctx, int n, const int* curr, const int* next, const int* prev, const double* rigidity,
                                   const double* curvature)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (n < 0 || (n > 0 && !rigidity)) return fail(ctx, IBK_ERR_INVALID, "bad beam arrays");
    if (int rc = check_nodes(ctx, n, { curr, next, prev }, "beams")) return rc;
    for (int k = 0; k < n; ++k)
        if (curr[k] == next[k] || curr[k] == prev[k]) return fail(ctx, IBK_ERR_INVALID, "beam repeats its master node"); // :1075-1076
    const int nn = force_node_bound(lv);
    std::vector<std::pair<int, int>> node_item;
    node_item.reserve(3 * (size_t)n);
    for (int k = 0; k < n; ++k)
    {
        node_item.push_back({ curr[k], 4 * k });
        node_item.push_back({ next[k], 4 * k + 1 });
        node_item.push_back({ prev[k], 4 * k + 2 });
    }
    std::vector<int> ptr, items;
    build_csr(nn, node_item, ptr, items);
    std::vector<double> curv((size_t)n * lv.ndim, 0.0);
    if (curvature) curv.assign(curvature, curvature + (size_t)n * lv.ndim);
    CK(force_upload(lv, ptr, &lv.force.beam_ptr));
    CK(force_upload(lv, items, &lv.force.beam_items));
    CK(force_upload(lv, std::vector<int>(curr, curr + n), &lv.force.beam_mastr));
    CK(force_upload(lv, std::vector<int>(next, next + n), &lv.force.beam_next));
    CK(force_upload(lv, std::vector<int>(prev, prev + n), &lv.force.beam_prev));
    CK(force_upload(lv, std::vector<double>(rigidity, rigidity + n), &lv.force.beam_rigidity));
    CK(force_upload(lv, curv, &lv.force.beam_curvature));
    return force_commit(ctx);
}

extern "C" int ibk_force_set_target_points(ibk_ctx* ctx, int n, const int* idx, const double* kappa, const double* eta,
                                           const double* X0)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (n < 0 || (n > 0 && (!kappa || !X0))) return fail(ctx, IBK_ERR_INVALID, "bad target point arrays");
    if (int rc = check_nodes(ctx, n, { idx }, "target points")) return rc;
    const int nn = force_node_bound(lv);
    std::vector<std::pair<int, int>> node_item;
    for (int k = 0; k < n; ++k) node_item.push_back({ idx[k], k });
    std::vector<int> ptr, items;
    build_csr(nn, node_item, ptr, items);
    std::vector<double> e((size_t)n, 0.0);
    if (eta) e.assign(eta, eta + n);
    CK(force_upload(lv, ptr, &lv.force.target_ptr));
    CK(force_upload(lv, items, &lv.force.target_items));
    CK(force_upload(lv, std::vector<double>(kappa, kappa + n), &lv.force.target_kappa));
    CK(force_upload(lv, e, &lv.force.target_eta));
    CK(force_upload(lv, std::vector<double>(X0, X0 + (size_t)n * lv.ndim), &lv.force.target_X0));
    return force_commit(ctx);
}

// Lagrangian index -> storage position, rebuilt after the storage order changed
static int refresh_pos_of_id(ibk_ctx* ctx)
{
    LevelState& lv = ctx->lv;
    const int bound = std::max(force_node_bound(lv), 1);
    if (lv.pos_valid && lv.pos_cap >= bound) return IBK_OK;
    if (lv.pos_cap < bound)
    {
        if (lv.pos_of_id) cudaFree(lv.pos_of_id);
        lv.pos_of_id = nullptr;
        CK(cudaMalloc(&lv.pos_of_id, sizeof(int) * (size_t)bound));
        lv.pos_cap = bound;
    }
    CK(launch_pos_of_id(ctx->L, lv.gid ? lv.gid : lv.lag, lv.n, lv.pos_of_id, bound));
    lv.pos_valid = true;
    return IBK_OK;
}

extern "C" int ibk_compute_lagrangian_force(ibk_ctx* ctx, int x_col, int u_col, int f_col)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    for (int c : { x_col, u_col, f_col })
        if (int rc = ensure_column(ctx, c)) return rc;
    if (f_col == x_col || f_col == u_col) return fail(ctx, IBK_ERR_INVALID, "the force column must differ from its inputs");
    if (lv.n == 0) return IBK_OK;
    if (lv.force.n_nodes != force_node_bound(lv) && (lv.force.spring_ptr || lv.force.beam_ptr || lv.force.target_ptr))
        return fail(ctx, IBK_ERR_STATE, "the force elements were set for a different marker numbering (set them after the markers)");
    if (int rc = refresh_pos_of_id(ctx)) return rc;
    if (!lv.d_missing) CK(cudaMalloc(&lv.d_missing, sizeof(int)));
    CK(cudaMemsetAsync(lv.d_missing, 0, sizeof(int), ctx->L.stream));
    CK(launch_lagrangian_force(ctx->L, lv.ndim, lv.force, column_of(lv, x_col), column_of(lv, u_col), column_of(lv, f_col), lv.stride,
                               lv.gid ? lv.gid : lv.lag, lv.pos_of_id, lv.n, lv.d_missing));
    if (lv.gid) // several processes: an element may reach a node that lives elsewhere (the reference's ghost nodes)
    {
        int missing = 0;
        CK(cudaMemcpyAsync(&missing, lv.d_missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
        if (missing > 0)
            return fail(ctx, IBK_ERR_STATE, "a force element reaches a node held by another process (structures must not straddle ranks)");
    }
    return IBK_OK;
}

extern "C" int ibk_markers_lincomb(ibk_ctx* ctx, int dst, double alpha, int a, double beta, int b)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    for (int c : { dst, a, b })
        if (int rc = ensure_column(ctx, c)) return rc;
    if (lv.n == 0) return IBK_OK;
    CK(launch_lincomb(ctx->L, column_of(lv, dst), alpha, column_of(lv, a), beta, column_of(lv, b), lv.stride, lv.n, lv.ndim));
    if (dst == 0) lv.binned = false; // the working positions moved: ibk_rebin before the next spread / interpolation
    return IBK_OK;
}

extern "C" int ibk_markers_scale_rows(ibk_ctx* ctx, int dst, int src, const double* h_ds)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    for (int c : { dst, src })
        if (int rc = ensure_column(ctx, c)) return rc;
    if (!h_ds && lv.n > 0) return fail(ctx, IBK_ERR_INVALID, "null weight array");
    if (lv.n == 0) return IBK_OK;
    CK(ctx->b_io[7].reserve(sizeof(double) * (size_t)lv.n));
    CK(cudaMemcpyAsync(ctx->b_io[7].p, h_ds, sizeof(double) * (size_t)lv.n, cudaMemcpyHostToDevice, ctx->L.stream));
    CK(launch_scale_rows(ctx->L, column_of(lv, dst), column_of(lv, src), lv.stride, lv.n, lv.ndim, ctx->b_io[7].as<double>(), lv.lag));
    CK(cudaStreamSynchronize(ctx->L.stream)); // h_ds is the caller's (pageable) memory
    if (dst == 0) lv.binned = false;
    return IBK_OK;
}

extern "C" int ibk_markers_zero_rows(ibk_ctx* ctx, int which, const int* lag_idx, int n)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (int rc = ensure_column(ctx, which)) return rc;
    if (n < 0 || (n > 0 && !lag_idx)) return fail(ctx, IBK_ERR_INVALID, "bad index list");
    if (n == 0 || lv.n == 0) return IBK_OK;
    if (int rc = refresh_pos_of_id(ctx)) return rc;
    CK(ctx->b_io[6].reserve(sizeof(int) * (size_t)n));
    CK(cudaMemcpyAsync(ctx->b_io[6].p, lag_idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    CK(launch_zero_rows(ctx->L, column_of(lv, which), lv.stride, lv.ndim, ctx->b_io[6].as<int>(), n, lv.pos_of_id,
                        std::max(force_node_bound(lv), 1)));
    CK(cudaStreamSynchronize(ctx->L.stream)); // lag_idx is the caller's (pageable) memory
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// halo
// ---------------------------------------------------------------------------------------------
static unsigned halo_blocks(const LevelState& lv, const PatchState& ps, int axis)
{
    // enough CTAs to cover the largest slab (a face of the array times the ghost width) a few times over
    long long face = 1;
    for (int d = 0; d < lv.ndim; ++d) face = std::max(face, (long long)ps.n[axis][(d + 1) % lv.ndim] * ps.n[axis][(d + 2) % lv.ndim]);
    long long nb = (face * 4 + 255) / 256;
    return (unsigned)std::min<long long>(std::max<long long>(nb, 1), 148 * 8);
}

extern "C" int ibk_spread_fold_walls(ibk_ctx* ctx);
extern "C" int ibk_halo_local(ibk_ctx* ctx, int which)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "halo plan missing");
    if (which < 0 || which > 1) return fail(ctx, IBK_ERR_INVALID, "which must be 0 (u: fill) or 1 (f: accumulate)");
    GRID_DEPS(which);
    if (which == 1)
        if (int rc = ibk_spread_fold_walls(ctx)) return rc;
    if (which == 1 && ex->halo.n_all_pairs > 0)
    {
        face_sync_kernel<<<dim3(64, ex->halo.n_all_pairs), 256, 0, ctx->L.stream>>>(ex->halo.d_all_pairs);
        ctx->L.launches++;
    }
    if (which == 0)
        CK(launch_region_items<0>(ctx, ex->halo, 0));
    else
        CK(launch_region_items<1>(ctx, ex->halo, 1));
    return IBK_OK;
}

static int region_args(ibk_ctx* ctx, int which, int patch, int axis, const int* lower, const int* upper, int* off, int* ext)
{
    if (int rc = check_patch_axis(ctx, which, patch, axis)) return rc;
    const LevelState& lv = ctx->lv;
    const PatchState& ps = lv.patches[patch];
    for (int d = 0; d < 3; ++d)
    {
        if (d < lv.ndim)
        {
            off[d] = lower[d] - (ps.lower[d] - lv.gcw[d]);
            ext[d] = upper[d] - lower[d] + 1;
            if (off[d] < 0 || ext[d] < 0 || off[d] + ext[d] > ps.n[axis][d]) return fail(ctx, IBK_ERR_INVALID, "region outside the array");
        }
        else
        {
            off[d] = 0;
            ext[d] = 1;
        }
    }
    return IBK_OK;
}
extern "C" int ibk_halo_pack(ibk_ctx* ctx, int which, int patch, int axis, const int* lower, const int* upper, double* d_buf)
{
    NEED_LEVEL();
    int off[3], ext[3];
    if (int rc = region_args(ctx, which, patch, axis, lower, upper, off, ext)) return rc;
    GRID_DEPS(which);
    PatchState& ps = ctx->lv.patches[patch];
    CK(launch_pack(ctx->L, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], ps.n[axis][1], off, ext, d_buf, ctx->lv.ndim));
    return IBK_OK;
}
extern "C" int ibk_halo_unpack(ibk_ctx* ctx, int which, int patch, int axis, const int* lower, const int* upper,
                               const double* d_buf, int mode)
{
    NEED_LEVEL();
    int off[3], ext[3];
    if (int rc = region_args(ctx, which, patch, axis, lower, upper, off, ext)) return rc;
    GRID_DEPS(which);
    PatchState& ps = ctx->lv.patches[patch];
    CK(launch_unpack(ctx->L, which == 0 ? ps.u[axis] : ps.f[axis], ps.pitch[axis], ps.n[axis][1], off, ext, d_buf, ctx->lv.ndim,
                     mode));
    return IBK_OK;
}
// One call per message: the items of a neighbour's buffer in order (same kernels as ibk_halo_pack / _unpack; the
// host side of a many-region exchange then costs one library call instead of one per region).
static int item_table(ibk_ctx* ctx, int which, int n_items, const int* patch, const int* axis, const int* lower, const int* upper,
                      const long long* buf_offset, const ItemTable** out)
{
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "level parameters missing");
    const int ndim = ctx->lv.ndim;
    unsigned long long key = 1469598103934665603ull; // FNV-1a over everything that defines the table
    auto mix = [&](long long v) {
        for (int b = 0; b < 8; ++b)
        {
            key ^= (unsigned long long)((v >> (8 * b)) & 0xff);
            key *= 1099511628211ull;
        }
    };
    mix(which);
    mix(n_items);
    for (int k = 0; k < n_items; ++k)
    {
        mix(patch[k]);
        mix(axis[k]);
        mix(buf_offset[k]);
        for (int d = 0; d < ndim; ++d)
        {
            mix(lower[(size_t)k * ndim + d]);
            mix(upper[(size_t)k * ndim + d]);
        }
    }
    for (const ItemTable& t : ex->item_tables)
        if (t.key == key && t.n == n_items)
        {
            *out = &t;
            return IBK_OK;
        }
    std::vector<HaloItem> h(n_items);
    ItemTable t;
    t.key = key;
    t.n = n_items;
    for (int k = 0; k < n_items; ++k)
    {
        int off[3], ext[3];
        if (int rc = region_args(ctx, which, patch[k], axis[k], lower + (size_t)k * ndim, upper + (size_t)k * ndim, off, ext)) return rc;
        PatchState& ps = ctx->lv.patches[patch[k]];
        h[k].ptr = which == 0 ? ps.u[axis[k]] : ps.f[axis[k]];
        h[k].pitch = ps.pitch[axis[k]];
        h[k].n1 = ps.n[axis[k]][1];
        h[k].count = 1;
        for (int d = 0; d < 3; ++d)
        {
            h[k].off[d] = off[d];
            h[k].ext[d] = ext[d];
            h[k].count *= ext[d];
        }
        h[k].buf_off = buf_offset[k];
        t.max_count = std::max(t.max_count, h[k].count);
        if (t.wave_start.empty()) t.wave_start.push_back(0);
        bool clash = false; // does the region overlap one of the current wave?
        for (int k2 = t.wave_start.back(); k2 < k; ++k2)
        {
            if (h[k2].ptr != h[k].ptr) continue;
            bool ov = true;
            for (int d = 0; d < 3; ++d) ov = ov && h[k].off[d] < h[k2].off[d] + h[k2].ext[d] && h[k2].off[d] < h[k].off[d] + h[k].ext[d];
            clash = clash || ov;
        }
        if (clash) t.wave_start.push_back(k);
    }
    t.wave_start.push_back(n_items);
    // CTAs: one numbering for the single-launch pack (block0 over all items) is not compatible with per-wave launches, so
    // the table is kept twice: [0, n) numbered over all items, [n, 2n) numbered per wave
    h.resize(2 * (size_t)n_items);
    for (int k = 0; k < n_items; ++k)
    {
        h[k].nblocks = region_blocks(h[k].ext);
        h[k].block0 = t.all_blocks;
        t.all_blocks += h[k].nblocks;
        h[n_items + k] = h[k];
    }
    for (size_t w = 0; w + 1 < t.wave_start.size(); ++w)
    {
        unsigned nb = 0;
        for (int k = t.wave_start[w]; k < t.wave_start[w + 1]; ++k)
        {
            h[n_items + k].block0 = nb;
            nb += h[n_items + k].nblocks;
        }
        t.wave_blocks.push_back(nb);
    }
    if (n_items > 0)
    {
        t.h_items.assign(h.begin(), h.begin() + n_items);
        CK(cudaMalloc(&t.d_items, sizeof(HaloItem) * 2 * (size_t)n_items));
        CK(cudaMemcpy(t.d_items, h.data(), sizeof(HaloItem) * 2 * (size_t)n_items, cudaMemcpyHostToDevice));
    }
    ex->item_tables.push_back(t);
    *out = &ex->item_tables.back();
    return IBK_OK;
}

// One call and (when the regions of an array are disjoint) one launch per message: the items of a neighbour's
// buffer in order.  The item table is kept on the device, keyed by its content.
extern "C" int ibk_halo_pack_many(ibk_ctx* ctx, int which, int n_items, const int* patch, const int* axis, const int* lower,
                                  const int* upper, const long long* buf_offset, double* d_buf)
{
    NEED_LEVEL();
    if (n_items < 0 || (n_items > 0 && (!patch || !axis || !lower || !upper || !buf_offset || !d_buf)))
        return fail(ctx, IBK_ERR_INVALID, "bad item arrays");
    if (which < 0 || which > 1) return fail(ctx, IBK_ERR_INVALID, "which must be 0 (u) or 1 (f)");
    if (n_items == 0) return IBK_OK;
    GRID_DEPS(which);
    const ItemTable* t = nullptr;
    if (int rc = item_table(ctx, which, n_items, patch, axis, lower, upper, buf_offset, &t)) return rc;
    CK(launch_halo_items(ctx->L, t->d_items, n_items, t->all_blocks, d_buf, 0));
    return IBK_OK;
}
extern "C" int ibk_halo_unpack_many(ibk_ctx* ctx, int which, int n_items, const int* patch, const int* axis, const int* lower,
                                    const int* upper, const long long* buf_offset, const double* d_buf, int mode)
{
    NEED_LEVEL();
    if (n_items < 0 || (n_items > 0 && (!patch || !axis || !lower || !upper || !buf_offset || !d_buf)))
        return fail(ctx, IBK_ERR_INVALID, "bad item arrays");
    if (which < 0 || which > 1 || mode < 0 || mode > 1) return fail(ctx, IBK_ERR_INVALID, "bad which / mode");
    if (n_items == 0) return IBK_OK;
    GRID_DEPS(which);
    const ItemTable* t = nullptr;
    if (int rc = item_table(ctx, which, n_items, patch, axis, lower, upper, buf_offset, &t)) return rc;
    if (mode == 1 && t->wave_start.size() > 2)
    {
        // overlapping regions: ONE launch over boxes that each add their buffer sources in list order
        ItemTable* tt = const_cast<ItemTable*>(t);
        if (!tt->d_add_ops || tt->add_buf != d_buf)
        {
            if (tt->d_add_ops) cudaFree(tt->d_add_ops);
            tt->d_add_ops = nullptr;
            struct Cell
            {
                double* dst;
                HBox box; // in the array coordinates of dst
                std::vector<int> src;
            };
            std::vector<Cell> cells;
            const std::vector<HaloItem>& H = tt->h_items;
            for (int k = 0; k < n_items; ++k)
            {
                HBox B;
                for (int d = 0; d < 3; ++d)
                {
                    B.lo[d] = H[k].off[d];
                    B.hi[d] = H[k].off[d] + H[k].ext[d] - 1;
                }
                std::vector<Cell> next;
                std::vector<HBox> uncovered{ B };
                for (Cell& c : cells)
                {
                    HBox J;
                    if (c.dst != H[k].ptr || !hbox_intersect(c.box, B, J))
                    {
                        next.push_back(c);
                        continue;
                    }
                    Cell both{ c.dst, J, c.src };
                    both.src.push_back(k);
                    next.push_back(both);
                    std::vector<HBox> rest;
                    hbox_minus(c.box, J, rest);
                    for (const HBox& R : rest) next.push_back(Cell{ c.dst, R, c.src });
                    std::vector<HBox> unc2;
                    for (const HBox& U : uncovered) hbox_minus(U, J, unc2);
                    uncovered.swap(unc2);
                }
                for (const HBox& U : uncovered) next.push_back(Cell{ H[k].ptr, U, std::vector<int>{ k } });
                cells.swap(next);
            }
            std::vector<RegionMulti> ops;
            unsigned nb = 0;
            for (const Cell& c : cells)
            {
                // (more sources than a RegionMulti holds: a second box over the same elements, later in the list)
                for (size_t s0 = 0; s0 < c.src.size(); s0 += REGION_MAXSRC)
                {
                    RegionMulti op;
                    std::memset(&op, 0, sizeof(op));
                    const HaloItem& first = H[c.src[0]];
                    op.dst = c.dst;
                    op.dst_pitch = first.pitch;
                    op.dst_n1 = first.n1;
                    for (int d = 0; d < 3; ++d)
                    {
                        op.dst_off[d] = c.box.lo[d];
                        op.ext[d] = c.box.hi[d] - c.box.lo[d] + 1;
                    }
                    for (size_t q = s0; q < std::min(c.src.size(), s0 + REGION_MAXSRC); ++q)
                    {
                        const HaloItem& it = H[c.src[q]];
                        const int cc = op.nsrc++;
                        op.src[cc] = d_buf + it.buf_off; // the item's region, dense in the buffer
                        op.src_pitch[cc] = it.ext[0];
                        op.src_n1[cc] = it.ext[1];
                        for (int d = 0; d < 3; ++d) op.src_off[cc][d] = c.box.lo[d] - it.off[d];
                    }
                    op.block0 = nb;
                    op.nblocks = region_blocks(op.ext);
                    nb += op.nblocks;
                    ops.push_back(op);
                }
            }
            // boxes over the same elements (a cell split for its many sources) must not run in one launch: rare enough to
            // fall back to the waves
            bool split = false;
            for (const Cell& c : cells) split = split || c.src.size() > (size_t)REGION_MAXSRC;
            if (!split && !ops.empty())
            {
                CK(cudaMalloc(&tt->d_add_ops, sizeof(RegionMulti) * ops.size()));
                CK(cudaMemcpyAsync(tt->d_add_ops, ops.data(), sizeof(RegionMulti) * ops.size(), cudaMemcpyHostToDevice, ctx->L.stream));
                CK(cudaStreamSynchronize(ctx->L.stream));
                tt->n_add_ops = (int)ops.size();
                tt->add_blocks = nb;
                tt->add_buf = d_buf;
            }
        }
        if (tt->d_add_ops)
        {
            region_items_kernel<1><<<tt->add_blocks, 256, 0, ctx->L.stream>>>(tt->d_add_ops, tt->n_add_ops);
            ctx->L.launches++;
            CK(cudaGetLastError());
            return IBK_OK;
        }
    }
    // one launch per wave of pairwise disjoint regions, the waves in list order (fixed order of the additions)
    for (size_t w = 0; w + 1 < t->wave_start.size(); ++w)
    {
        const int k0 = t->wave_start[w], k1 = t->wave_start[w + 1];
        if (k1 > k0) CK(launch_halo_items(ctx->L, t->d_items + n_items + k0, k1 - k0, t->wave_blocks[w], const_cast<double*>(d_buf), mode == 0 ? 1 : 2));
    }
    return IBK_OK;
}
// ---------------------------------------------------------------------------------------------
// spreadForce / interpolateVelocity
// ---------------------------------------------------------------------------------------------
static int face_layers(ibk_ctx* ctx, int mode)
{
    LevelState& lv = ctx->lv;
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "halo plan missing");
    (void)lv;
    if (ex->halo.n_face_jobs > 0)
    {
        face_layers_kernel<<<dim3(std::min(ex->halo.face_job_blocks, 256u), ex->halo.n_face_jobs), 256, 0, ctx->L.stream>>>(ex->halo.d_face_jobs, mode);
        ctx->L.launches++;
    }
    CK(cudaGetLastError());
    return IBK_OK;
}

extern "C" int ibk_spread_begin(ibk_ctx* ctx)
{
    NEED_LEVEL();
    GRID_DEPS(1);
    LevelState& lv = ctx->lv;
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "halo plan missing");
    // the ghost regions of f receive fresh spread values only (LDataManager.cpp:594)
    CK(launch_region_items<2>(ctx, ex->halo, 2));
    ex->halo.walls_folded = false;
    return face_layers(ctx, 0);
}
extern "C" int ibk_spread_end(ibk_ctx* ctx)
{
    NEED_LEVEL();
    GRID_DEPS(1);
    return face_layers(ctx, 1);
}

// Robin coefficients of the physical boundaries, [ndim][2 sides][ndim components] each (a u + b du/dn = g; the fold-back
// is the homogeneous adjoint and does not use g).  From then on every spread with halo handling folds what it put into
// the ghost cells outside the domain back into the interior before the ghost accumulation.  Null pointers switch it off.
extern "C" int ibk_level_set_wall_bc(ibk_ctx* ctx, const double* acoef, const double* bcoef)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "halo plan missing");
    HaloPlan& hp = ex->halo;
    cudaStreamSynchronize(ctx->L.stream);
    if (hp.d_wall_jobs) cudaFree(hp.d_wall_jobs);
    hp.d_wall_jobs = nullptr;
    hp.n_wall_jobs = 0;
    if (!acoef || !bcoef) return IBK_OK;
    const int ndim = lv.ndim;
    int nwall = 0;
    for (int d = 0; d < ndim; ++d) nwall += lv.periodic[d] ? 0 : 1;
    if (nwall > 1)
        return fail(ctx, IBK_ERR_INVALID,
                    "physical boundaries in more than one dimension: the co-dimension two / three extrapolations "
                    "(CartSideRobinPhysBdryOp.cpp:586-598) are not built");
    std::vector<WallJob> jobs;
    for (const PatchState& ps : lv.patches)
        for (int d = 0; d < ndim; ++d)
        {
            if (lv.periodic[d]) continue;
            for (int side = 0; side < 2; ++side)
            {
                const bool touches = side == 0 ? ps.lower[d] == lv.domain_lower[d] : ps.upper[d] == lv.domain_upper[d];
                if (!touches) continue;
                for (int comp = 0; comp < ndim; ++comp)
                {
                    WallJob J;
                    std::memset(&J, 0, sizeof(J));
                    J.ptr = ps.f[comp];
                    J.stride[0] = 1;
                    J.stride[1] = ps.pitch[comp];
                    J.stride[2] = ps.pitch[comp] * ps.n[comp][1];
                    for (int e = 0; e < 3; ++e) J.n[e] = ps.n[comp][e];
                    J.dim = d;
                    J.side = side;
                    J.normal = comp == d ? 1 : 0;
                    J.gcw = lv.gcw[d];
                    const int ncell = ps.upper[d] - ps.lower[d] + 1;
                    // array index = index - (lower - gcw): boundary face (normal) = lower / upper + 1, first interior cell
                    // (transverse) = lower / upper
                    if (J.normal) J.ib = side == 0 ? lv.gcw[d] : lv.gcw[d] + ncell;
                    else J.ib = side == 0 ? lv.gcw[d] : lv.gcw[d] + ncell - 1;
                    J.a = acoef[(d * 2 + side) * ndim + comp];
                    J.b = bcoef[(d * 2 + side) * ndim + comp];
                    J.h = lv.dx[d];
                    jobs.push_back(J);
                    const long long cols = (long long)J.n[(d + 1) % 3] * J.n[(d + 2) % 3];
                    hp.wall_blocks = std::max<unsigned>(hp.wall_blocks, (unsigned)std::min<long long>(148 * 8, (cols + 255) / 256));
                }
            }
        }
    if (jobs.empty()) return IBK_OK;
    CK(cudaMalloc(&hp.d_wall_jobs, sizeof(WallJob) * jobs.size()));
    CK(cudaMemcpy(hp.d_wall_jobs, jobs.data(), sizeof(WallJob) * jobs.size(), cudaMemcpyHostToDevice));
    hp.n_wall_jobs = (int)jobs.size();
    return IBK_OK;
}
// The fold-back itself: once per spread (after the tiles that can reach a ghost cell, before any ghost value is packed or
// accumulated).  ibk_spread_force and ibk_halo_local(f) / ibk_halo_accumulate_post call it themselves.
extern "C" int ibk_spread_fold_walls(ibk_ctx* ctx)
{
    NEED_LEVEL();
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "halo plan missing");
    HaloPlan& hp = ex->halo;
    if (hp.n_wall_jobs <= 0 || hp.walls_folded) return IBK_OK;
    GRID_DEPS(1);
    wall_fold_kernel<<<dim3(hp.wall_blocks, hp.n_wall_jobs), 256, 0, ctx->L.stream>>>(hp.d_wall_jobs);
    ctx->L.launches++;
    hp.walls_folded = true;
    CK(cudaGetLastError());
    return IBK_OK;
}

static int level_op(ibk_ctx* ctx, int op, const char* fcn, int halo, int part = 0)
{
    NEED_LEVEL();
    GRID_DEPS(op == 1 ? 1 : 0);
    LevelState& lv = ctx->lv;
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "markers are not binned: call ibk_rebin first");
    const int kernel = ibk_kernel_from_string(fcn);
    if (kernel < 0) return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, std::string("unknown kernel function ") + (fcn ? fcn : "(null)"));
    if (kernel == IBK_USER_DEFINED)
        return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, "USER_DEFINED is a host callback: it is served at the patch seams (ibk_side_*_host ...), not on the resident level");
    const int min_ghosts = ibk_get_minimum_ghost_width(fcn);
    for (int d = 0; d < lv.ndim; ++d)
        if (lv.gcw[d] < min_ghosts) return fail(ctx, IBK_ERR_GHOST_WIDTH, "insufficient ghost cells for the kernel function");
    LevelExtra* ex = extra_of(ctx, false);
    if (!ex) return fail(ctx, IBK_ERR_STATE, "level parameters missing");
    const int which_ev = op == 1 ? 0 : 1;
    if (op == 0 && halo)
        if (int rc = ibk_halo_local(ctx, 0)) return rc;
    if (op == 1 && halo)
        if (int rc = ibk_spread_begin(ctx)) return rc;
    if (ctx->timing) CK(cudaEventRecord(ctx->ev[which_ev][0], ctx->L.stream));
    MarkerView mv;
    mv.X = lv.X;
    mv.Xraw = nullptr;
    mv.x_stride = lv.stride;
    mv.V = op == 0 ? lv.U : lv.F;
    mv.v_cstride = lv.stride;
    mv.v_istride = 1;
    mv.src = nullptr;
    mv.clip_free = true; // owners only, ghost width checked above: no stencil leaves the arrays
    if (part < 0 || part > 2) return fail(ctx, IBK_ERR_INVALID, "part must be 0 (all), 1 (interior tiles) or 2 (boundary tiles)");
    mv.part = part;
    for (size_t p = 0; p < lv.patches.size(); ++p)
    {
        if (part)
        {
            // Interior tiles: neither the block of the spread (tile -+ M points, + the TMA alignment column) nor the
            // staged box of the interpolation (tile -+ M + 2) reaches a ghost cell or a boundary face of the patch,
            // i.e. anything the halo exchange reads or writes.  pp = index - lower + G.
            const PatchState& ps = lv.patches[p];
            const int reach = (ibk_get_stencil_size(fcn) + 1) / 2 + 2; // kernel reach M + the spare columns of the TMA boxes
            for (int d = 0; d < 3; ++d)
            {
                if (d >= lv.ndim)
                {
                    mv.sel_lo[d] = 0;
                    mv.sel_hi[d] = 0;
                    continue;
                }
                const int n = ps.upper[d] - ps.lower[d] + 1;
                mv.sel_lo[d] = (lv.G + reach) / 16 + 1;     // 16 t - reach > G: clear of the low ghosts and the boundary face
                mv.sel_hi[d] = (lv.G + n - 17 - reach) / 16; // 16 t + 16 + reach <= G + n - 1: clear of the high face and ghosts
                if (op == 1 && lv.ndim == 3 && d < 2)
                {
                    // the march kernel works on columns of 2 x 2 tiles: a column cut by the selection would be marched
                    // twice, half empty each time (measured: + 0.4 ms on the C5 shard).  Whole columns go to one part.
                    mv.sel_lo[d] += mv.sel_lo[d] & 1;
                    mv.sel_hi[d] -= !(mv.sel_hi[d] & 1);
                }
            }
        }
        std::string err;
        cudaError_t e = (op == 0) ? launch_interp(ctx->L, kernel, ex->tp[2 * p + 0], lv.bins, mv, err) :
                                    launch_spread(ctx->L, kernel, ex->tp[2 * p + 1], lv.bins, mv, err);
        if (e != cudaSuccess) return cuda_fail(ctx, e, err.empty() ? "tile kernel launch" : err.c_str());
    }
    if (ctx->timing)
    {
        CK(cudaEventRecord(ctx->ev[which_ev][1], ctx->L.stream));
        ctx->ev_valid[which_ev] = true;
    }
    if (op == 1 && halo)
    {
        if (int rc = ibk_halo_local(ctx, 1)) return rc;
        if (int rc = ibk_spread_end(ctx)) return rc;
    }
    return IBK_OK;
}

extern "C" int ibk_spread_force(ibk_ctx* ctx, const char* spread_fcn, int accumulate_halo)
{
    return level_op(ctx, 1, spread_fcn, accumulate_halo);
}
extern "C" int ibk_interpolate_velocity(ibk_ctx* ctx, const char* interp_fcn, int fill_halo)
{
    return level_op(ctx, 0, interp_fcn, fill_halo);
}

extern "C" int ibk_spread_force_part(ibk_ctx* ctx, const char* spread_fcn, int part)
{
    return level_op(ctx, 1, spread_fcn, 0, part);
}
extern "C" int ibk_interpolate_velocity_part(ibk_ctx* ctx, const char* interp_fcn, int part)
{
    return level_op(ctx, 0, interp_fcn, 0, part);
}

extern "C" int ibk_count_touched_dofs(ibk_ctx* ctx, const char* kernel_fcn, long long* touched)
{
    NEED_LEVEL();
    LevelState& lv = ctx->lv;
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "markers are not binned: call ibk_rebin first");
    const int kernel = ibk_kernel_from_string(kernel_fcn);
    if (kernel < 0 || !touched) return fail(ctx, IBK_ERR_UNKNOWN_KERNEL, "unknown kernel function");
    LevelExtra* ex = extra_of(ctx, false);
    // spread a field of ones into scratch copies of f and count the nonzeros
    DevBuf ones, cnt;
    CK(ones.reserve(sizeof(double) * (size_t)lv.stride * lv.ndim));
    CK(cnt.reserve(sizeof(unsigned long long)));
    CK(launch_fill(ctx->L, ones.as<double>(), (size_t)lv.stride * lv.ndim, 1.0));
    CK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), ctx->L.stream));
    MarkerView mv;
    mv.X = lv.X;
    mv.Xraw = nullptr;
    mv.x_stride = lv.stride;
    mv.V = ones.as<double>();
    mv.v_cstride = lv.stride;
    mv.v_istride = 1;
    mv.src = nullptr;
    int rc = IBK_OK;
    for (size_t p = 0; p < lv.patches.size() && rc == IBK_OK; ++p)
    {
        const PatchState& ps = lv.patches[p];
        TileParams tp = ex->tp[2 * p + 1];
        DevBuf scratch[3];
        for (int a = 0; a < lv.ndim; ++a)
        {
            CK(scratch[a].reserve(sizeof(double) * ps.elems[a]));
            CK(cudaMemsetAsync(scratch[a].p, 0, sizeof(double) * ps.elems[a], ctx->L.stream));
            tp.comp[a].ptr = scratch[a].as<double>();
        }
        std::string err;
        cudaError_t e = launch_spread(ctx->L, kernel, tp, lv.bins, mv, err);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "count_touched spread");
        for (int a = 0; a < lv.ndim && rc == IBK_OK; ++a)
        {
            count_nonzero_kernel<<<148 * 8, 256, 0, ctx->L.stream>>>(scratch[a].as<double>(), ps.pitch[a], ps.n[a][0],
                                                                     (long long)ps.n[a][1] * ps.n[a][2],
                                                                     cnt.as<unsigned long long>());
            ctx->L.launches++;
        }
        cudaStreamSynchronize(ctx->L.stream);
        for (int a = 0; a < 3; ++a) scratch[a].release();
    }
    unsigned long long h = 0;
    if (rc == IBK_OK)
    {
        CK(cudaMemcpyAsync(&h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
        *touched = (long long)h;
    }
    ones.release();
    cnt.release();
    return rc;
}
