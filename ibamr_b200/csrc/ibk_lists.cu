// ibk_lists.cu -- the index sets of LIndexSetData::cacheLocalIndices (ibtk/src/lagrangian/LIndexSetData.cpp:53-141) emitted
// from the device-resident binning products, for the bit-exact comparison north_star asks for.
//
// For a patch the reference lists every marker whose cell -- or a periodic image of it -- lies in the patch's GHOST box
// (patch grown by the ghost width), in the order of its cell-indexed container (cell k-j-i inside the ghost box, then
// Lagrangian index, LDataManager.cpp:1505, 2897-2911), each with the periodic shift of the image (:89-101) and the
// interior / ghost distinction (:104, 117-137).  The hot path here does not need these lists (owner-only spreading, DESIGN
// section 4); they are products for callers and for the parity tests.  The device holds the markers THIS process owns:
// with several processes a patch's ghost entries that live on other ranks are not listed (the lists are complete when all
// patches of the level are local).
//
// emit: one thread per marker tests the (up to 27) periodic images against the ghost box and appends
//       key = (cell number in the ghost box) << 32 | Lagrangian index, value = image code | interior flag;
// sort: the library's own radix sort; unpack on the host.
#include <cuda_runtime.h>

#include <vector>

#include "ibk_ctx.h"

namespace ibk
{
int fail(ibk_ctx* ctx, int code, const std::string& msg);
int cuda_fail(ibk_ctx* ctx, cudaError_t e, const char* what);

struct ListGeom
{
    int ndim;
    int glo[3], ghi[3], plo[3], phi[3]; // ghost box, patch box
    int ncells[3], periodic[3];
};

__global__ void patch_list_emit_kernel(ListGeom g, const int* __restrict__ cells, const uint32_t* __restrict__ lag, int n,
                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int capacity, int* __restrict__ counter)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3] = { 0, 0, 0 };
    for (int d = 0; d < g.ndim; ++d) c[d] = cells[(size_t)i * g.ndim + d];
    int omin[3] = { 0, 0, 0 }, omax[3] = { 0, 0, 0 };
    for (int d = 0; d < g.ndim; ++d)
        if (g.periodic[d])
        {
            omin[d] = -1;
            omax[d] = 1;
        }
    const long long gn0 = g.ghi[0] - g.glo[0] + 1, gn1 = g.ghi[1] - g.glo[1] + 1;
    for (int o2 = omin[2]; o2 <= omax[2]; ++o2)
        for (int o1 = omin[1]; o1 <= omax[1]; ++o1)
            for (int o0 = omin[0]; o0 <= omax[0]; ++o0)
            {
                const int o[3] = { o0, o1, o2 };
                int q[3] = { 0, 0, 0 };
                bool in_ghost = true, in_patch = true;
                for (int d = 0; d < g.ndim; ++d)
                {
                    q[d] = c[d] + o[d] * g.ncells[d];
                    in_ghost = in_ghost && q[d] >= g.glo[d] && q[d] <= g.ghi[d];
                    in_patch = in_patch && q[d] >= g.plo[d] && q[d] <= g.phi[d];
                }
                if (!in_ghost) continue;
                const int slot = atomicAdd(counter, 1);
                if (slot >= capacity) continue; // (counting pass: capacity 0)
                const long long cellkey = ((long long)(q[2] - g.glo[2]) * gn1 + (q[1] - g.glo[1])) * gn0 + (q[0] - g.glo[0]);
                keys[slot] = ((uint64_t)cellkey << 32) | (uint64_t)lag[i];
                vals[slot] = (uint32_t)((o0 + 1) + 3 * (o1 + 1) + 9 * (o2 + 1)) | (in_patch ? 32u : 0u);
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

// n_entries: in = capacity of the output arrays, out = number of entries of the patch's list.  With null arrays (or a
// capacity that is too small) only the count is returned (IBK_OK).  h_lag_idx[n], h_shift[n][ndim], h_interior[n].
extern "C" int ibk_bin_get_patch_lists(ibk_ctx* ctx, int patch, int* n_entries, int* h_lag_idx, double* h_shift, int* h_interior)
{
    if (!ctx || !n_entries) return IBK_ERR_INVALID;
    LevelState& lv = ctx->lv;
    if (!lv.valid) return fail(ctx, IBK_ERR_STATE, "no level registered (ibk_level_create)");
    if (!lv.binned) return fail(ctx, IBK_ERR_STATE, "ibk_rebin has not run");
    if (patch < 0 || patch >= (int)lv.patches.size()) return fail(ctx, IBK_ERR_INVALID, "bad patch number");
    const int ndim = lv.ndim, n = lv.n;
    const int capacity = *n_entries;
    *n_entries = 0;
    if (n == 0) return IBK_OK;
    ListGeom g;
    g.ndim = ndim;
    long long gcells = 1;
    for (int d = 0; d < 3; ++d)
    {
        const PatchState& ps = lv.patches[patch];
        g.plo[d] = d < ndim ? ps.lower[d] : 0;
        g.phi[d] = d < ndim ? ps.upper[d] : 0;
        g.glo[d] = g.plo[d] - (d < ndim ? lv.gcw[d] : 0);
        g.ghi[d] = g.phi[d] + (d < ndim ? lv.gcw[d] : 0);
        g.ncells[d] = d < ndim ? lv.domain_upper[d] - lv.domain_lower[d] + 1 : 1;
        g.periodic[d] = d < ndim ? lv.periodic[d] : 0;
        gcells *= (long long)(g.ghi[d] - g.glo[d] + 1);
    }
    if (gcells >= (1ll << 31)) return fail(ctx, IBK_ERR_INVALID, "ghost box too large for the list keys");
    DevBuf cnt;
    CK(cnt.reserve(sizeof(int)));
    const int T = 256, nb = (n + T - 1) / T;
    // counting pass
    CK(cudaMemsetAsync(cnt.p, 0, sizeof(int), ctx->L.stream));
    patch_list_emit_kernel<<<nb, T, 0, ctx->L.stream>>>(g, lv.cells, lv.lag_prev, n, nullptr, nullptr, 0, cnt.as<int>());
    ctx->L.launches++;
    int total = 0;
    CK(cudaMemcpyAsync(&total, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    *n_entries = total;
    if (total == 0 || !h_lag_idx || !h_shift || !h_interior || capacity < total)
    {
        cnt.release();
        return IBK_OK;
    }
    DevBuf k[2], v[2], tmp;
    for (int b = 0; b < 2; ++b)
    {
        CK(k[b].reserve(sizeof(uint64_t) * (size_t)total));
        CK(v[b].reserve(sizeof(uint32_t) * (size_t)total));
    }
    CK(tmp.reserve(radix_sort_temp_bytes(total)));
    CK(cudaMemsetAsync(cnt.p, 0, sizeof(int), ctx->L.stream));
    patch_list_emit_kernel<<<nb, T, 0, ctx->L.stream>>>(g, lv.cells, lv.lag_prev, n, k[0].as<uint64_t>(), v[0].as<uint32_t>(), total, cnt.as<int>());
    ctx->L.launches++;
    int cell_bits = 1;
    while ((1ll << cell_bits) < gcells) ++cell_bits;
    const int which = radix_sort_pairs(k[0].as<uint64_t>(), v[0].as<uint32_t>(), k[1].as<uint64_t>(), v[1].as<uint32_t>(), total, 0,
                                       ((32 + cell_bits + 7) / 8) * 8, tmp.p, ctx->L.stream, &ctx->L.launches);
    std::vector<uint64_t> hk(total);
    std::vector<uint32_t> hv(total);
    CK(cudaMemcpyAsync(hk.data(), k[which].p, sizeof(uint64_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(hv.data(), v[which].p, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int e = 0; e < total; ++e)
    {
        h_lag_idx[e] = (int)(hk[e] & 0xffffffffull);
        const int code = (int)(hv[e] & 31u);
        const int o[3] = { code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1 };
        // LIndexSetData.cpp:89-101: the shift of the image, (cells of the offset) * dx
        for (int d = 0; d < ndim; ++d) h_shift[(size_t)e * ndim + d] = (double)(o[d] * g.ncells[d]) * lv.dx[d];
        h_interior[e] = (hv[e] & 32u) ? 1 : 0;
    }
    for (int b = 0; b < 2; ++b)
    {
        k[b].release();
        v[b].release();
    }
    tmp.release();
    cnt.release();
    return IBK_OK;
}
