// ibk_bin.cu -- marker -> (patch, cell) binning on the device (LIndexSetData / LDataManager
// redistribution role).
//
//   wrap_positions   LDataManager::beginDataRedistribution, ibtk/src/lagrangian/LDataManager.cpp:1397-1421
//   bin_keys         cell = IndexUtilities::getCellIndex(X, grid_geom, ratio)  (:1475,
//                    ibtk/include/ibtk/private/IndexUtilities-inl.h:62-73), owner = patch whose box
//                    contains the cell (:1476); key = (patch brick, cell in brick, Lagrangian index)
//   radix sort       ibk_sort.cu (replaces the per-cell std::sort by Lagrangian index, :1502-1508,
//                    and the patch/cell running numbering of computeNodeDistribution, :2897-2911)
//   brick_offsets    segment offsets per 4^ndim-cell brick (replaces the IndexData<LSet> containers)
//
// Sort key (64 bit): [ brick id | cell in brick (2 bits per dim) | tie (Lagrangian index) ].
// Bricks are numbered tile-major (a tile = 4^ndim bricks = 16^ndim cells) so that the markers of
// one interpolation tile are contiguous and every brick is one contiguous segment.
#include <algorithm>
#include <cstdio>

#include "ibk_engine.h"

namespace ibk
{
struct BinParams
{
    CellGeom cg;
    int n_patches;
    int key_tie_bits;
    int cshift;
};

void fill_patch_bin(PatchBin& pb, int ndim, const int* lower, const int* upper, const int* accept_lo, const int* accept_hi,
                    int G, int brick_base)
{
    pb.ndim = ndim;
    pb.G = G;
    pb.brick_base = brick_base;
    int ntiles = 1;
    for (int d = 0; d < 3; ++d)
    {
        if (d < ndim)
        {
            pb.lower[d] = lower[d];
            pb.upper[d] = upper[d];
            pb.accept_lo[d] = accept_lo[d];
            pb.accept_hi[d] = accept_hi[d];
            const int ncc = upper[d] - lower[d] + 1 + 2 * G;
            const int nb = (ncc + BRICK - 1) / BRICK;
            pb.nt[d] = (nb + TILE_BRICKS - 1) / TILE_BRICKS;
            pb.nb[d] = pb.nt[d] * TILE_BRICKS;
            ntiles *= pb.nt[d];
        }
        else
        {
            pb.lower[d] = pb.upper[d] = pb.accept_lo[d] = pb.accept_hi[d] = 0;
            pb.nb[d] = pb.nt[d] = 1;
        }
    }
    pb.nbricks = ntiles * (ndim == 3 ? 64 : 16);
}

__global__ void wrap_positions_kernel(DomainGeom dg, double* __restrict__ X, long long stride, int n, int* __restrict__ escaped,
                                      int check_only)
{
    const double TOL = 1.4901161193847656e-08; // sqrt(DBL_EPSILON), LDataManager.cpp:150
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int d = 0; d < dg.ndim; ++d)
    {
        double x = X[d * stride + i];
        const double lo = dg.x_lower[d], hi = dg.x_upper[d];
        if (dg.periodic[d])
        {
            const double L = __dsub_rn(hi, lo);
            int guard = 0;
            while (x < lo && guard++ < 64) x = __dadd_rn(x, L);
            while (x >= hi && guard++ < 64) x = __dsub_rn(x, L);
        }
        else
        {
            if (x < lo || x > hi) atomicAdd(escaped, 1);
            x = fmax(x, lo);
            x = fmin(x, __dsub_rn(hi, __dmul_rn(__dsub_rn(hi, lo), TOL)));
        }
        if (!check_only) X[d * stride + i] = x;
    }
}

__global__ void bin_keys_kernel(BinParams bp,
                                const PatchBin* __restrict__ patches,
                                const double* __restrict__ X,
                                long long stride,
                                const uint32_t* __restrict__ tie,
                                int n,
                                uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ vals,
                                int* __restrict__ cells_out,
                                int* __restrict__ owner_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ndim = bp.cg.ndim;
    int c[3] = { 0, 0, 0 };
    for (int d = 0; d < ndim; ++d)
    {
        const double x = X[d * stride + i];
        if (bp.cg.two_branch)
            c[d] = cell_index_1d(x, bp.cg.x_lower[d], bp.cg.x_upper[d], bp.cg.dx[d], bp.cg.ilower[d], bp.cg.iupper[d]);
        else
            c[d] = bp.cg.ilower[d] + (int)floor(__ddiv_rn(__dsub_rn(x, bp.cg.x_lower[d]), bp.cg.dx[d]));
    }
    int owner = -1;
    for (int p = 0; p < bp.n_patches; ++p)
    {
        const PatchBin& pb = patches[p];
        bool in = true;
        for (int d = 0; d < ndim; ++d) in = in && c[d] >= pb.accept_lo[d] && c[d] <= pb.accept_hi[d];
        if (in)
        {
            owner = p;
            break;
        }
    }
    const uint32_t t = tie ? tie[i] : (uint32_t)i;
    uint64_t bkey = ~0ull >> bp.key_tie_bits; // discard bucket: all ones above the tie bits
    if (owner >= 0)
    {
        const PatchBin& pb = patches[owner];
        const int cx = c[0] - pb.lower[0] + pb.G;
        const int cy = c[1] - pb.lower[1] + pb.G;
        if (ndim == 3)
        {
            const int cz = c[2] - pb.lower[2] + pb.G;
            const int brick = pb.brick_base + brick_id_3d(cx >> 2, cy >> 2, cz >> 2, pb.nt);
            bkey = ((uint64_t)brick << 6) | (uint64_t)(((cz & 3) << 4) | ((cy & 3) << 2) | (cx & 3));
        }
        else
        {
            const int brick = pb.brick_base + brick_id_2d(cx >> 2, cy >> 2, pb.nt);
            bkey = ((uint64_t)brick << 4) | (uint64_t)(((cy & 3) << 2) | (cx & 3));
        }
    }
    keys[i] = (bkey << bp.key_tie_bits) | (uint64_t)t;
    vals[i] = (uint32_t)i;
    if (cells_out)
        for (int d = 0; d < ndim; ++d) cells_out[(size_t)ndim * i + d] = c[d];
    if (owner_out) owner_out[i] = owner;
}

// ids of the bricks with more than `thresh` markers, appended in no particular order
__global__ void dense_bricks_kernel(const int* __restrict__ brick_start, int total_bricks, int thresh, int* __restrict__ list,
                                    int capacity, int* __restrict__ count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_bricks) return;
    if (brick_start[b + 1] - brick_start[b] > thresh)
    {
        const int pos = atomicAdd(count, 1);
        if (pos < capacity) list[pos] = b;
    }
}

// brick_start[b] = first sorted position whose brick id is >= b; brick_start[total] = n_active.
__global__ void brick_offsets_kernel(const uint64_t* __restrict__ keys, int n, int shift, int total_bricks,
                                     int* __restrict__ brick_start)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    // brick id of entry i, clamped to total_bricks for discarded entries and for the end sentinel
    long long cur = total_bricks, prev = -1;
    if (i < n)
    {
        const unsigned long long b = keys[i] >> shift;
        cur = b < (unsigned long long)total_bricks ? (long long)b : total_bricks;
    }
    if (i > 0)
    {
        const unsigned long long b = keys[i - 1] >> shift;
        prev = b < (unsigned long long)total_bricks ? (long long)b : total_bricks;
    }
    for (long long b = prev + 1; b <= cur; ++b) brick_start[b] = i;
}

__global__ void gather_columns_kernel(const double* __restrict__ in, long long in_stride, double* __restrict__ out,
                                      long long out_stride, const uint32_t* __restrict__ perm, int n, int ncols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = perm[i];
    for (int c = 0; c < ncols; ++c) out[c * out_stride + i] = in[c * in_stride + s];
}

__global__ void gather_column_sets_kernel(GatherSets sets, long long stride, const uint32_t* __restrict__ perm, int n, int ncols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = perm[i];
    double v[6][3];
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (k < sets.nsets)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < ncols) v[k][c] = sets.in[k][c * stride + s];
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (k < sets.nsets)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < ncols) sets.out[k][c * stride + i] = v[k][c];
}

// The radix sort orders by (brick, cell) only; the markers of one cell are then put into ascending tie (Lagrangian index)
// order here (LDataManager.cpp:1505): the thread of a run's first element sorts the run by insertion.
__global__ void sort_ties_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int n, int tie_bits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t hi = keys[i] >> tie_bits;
    if (i > 0 && (keys[i - 1] >> tie_bits) == hi) return; // not the first of its run
    int e = i + 1;
    while (e < n && (keys[e] >> tie_bits) == hi) ++e;
    for (int a = i + 1; a < e; ++a)
    {
        const uint64_t k = keys[a];
        const uint32_t v = vals[a];
        int b = a - 1;
        while (b >= i && keys[b] > k)
        {
            keys[b + 1] = keys[b];
            vals[b + 1] = vals[b];
            --b;
        }
        keys[b + 1] = k;
        vals[b + 1] = v;
    }
}

__global__ void scatter_columns_kernel(const double* __restrict__ in, long long in_stride, double* __restrict__ out,
                                       long long out_stride, const uint32_t* __restrict__ perm, int n, int ncols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = perm[i];
    for (int c = 0; c < ncols; ++c) out[c * out_stride + s] = in[c * in_stride + i];
}

__global__ void extract_low_kernel(const uint64_t* __restrict__ keys, uint32_t* __restrict__ out, int n, uint64_t mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(keys[i] & mask);
}

static int bits_for(unsigned long long v)
{
    int b = 0;
    while (v)
    {
        ++b;
        v >>= 1;
    }
    return b > 0 ? b : 1;
}

cudaError_t bins_reserve(Bins& b, int n_entries, int total_bricks)
{
    cudaError_t e = cudaSuccess;
    if (n_entries > b.capacity)
    {
        const int cap = std::max(n_entries, 1024);
        for (int k = 0; k < 2; ++k)
        {
            if (b.keys[k]) cudaFree(b.keys[k]);
            if (b.vals[k]) cudaFree(b.vals[k]);
            if ((e = cudaMalloc(&b.keys[k], sizeof(uint64_t) * (size_t)cap)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&b.vals[k], sizeof(uint32_t) * (size_t)cap)) != cudaSuccess) return e;
        }
        if (b.sort_temp) cudaFree(b.sort_temp);
        b.sort_temp_bytes = radix_sort_temp_bytes(cap);
        if ((e = cudaMalloc(&b.sort_temp, b.sort_temp_bytes)) != cudaSuccess) return e;
        if (b.exc_flags) cudaFree(b.exc_flags);
        const size_t words = (size_t)cap / 4 + 1;
        if ((e = cudaMalloc(&b.exc_flags, sizeof(unsigned) * words)) != cudaSuccess) return e;
        if ((e = cudaMemset(b.exc_flags, 0, sizeof(unsigned) * words)) != cudaSuccess) return e;
        if (!b.exc_count)
        {
            if ((e = cudaMalloc(&b.exc_count, sizeof(int))) != cudaSuccess) return e;
            if ((e = cudaMemset(b.exc_count, 0, sizeof(int))) != cudaSuccess) return e;
        }
        b.capacity = cap;
    }
    if (total_bricks + 1 > b.brick_capacity)
    {
        if (b.brick_start) cudaFree(b.brick_start);
        if ((e = cudaMalloc(&b.brick_start, sizeof(int) * (size_t)(total_bricks + 1))) != cudaSuccess) return e;
        b.brick_capacity = total_bricks + 1;
        // a march tile holds at least one marker tile (16 bricks in 2D, 64 in 3D)
        if (b.march_sync) cudaFree(b.march_sync);
        b.march_sync_capacity = ((size_t)total_bricks / 16 + 8) * IBK_MAX_COMP + 1;
        if ((e = cudaMalloc(&b.march_sync, sizeof(int) * b.march_sync_capacity)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void bins_free(Bins& b)
{
    for (int k = 0; k < 2; ++k)
    {
        if (b.keys[k]) cudaFree(b.keys[k]);
        if (b.vals[k]) cudaFree(b.vals[k]);
        b.keys[k] = nullptr;
        b.vals[k] = nullptr;
    }
    if (b.sort_temp) cudaFree(b.sort_temp);
    if (b.brick_start) cudaFree(b.brick_start);
    if (b.dense_list) cudaFree(b.dense_list);
    if (b.dense_count) cudaFree(b.dense_count);
    if (b.march_sync) cudaFree(b.march_sync);
    if (b.exc_flags) cudaFree(b.exc_flags);
    if (b.exc_count) cudaFree(b.exc_count);
    b = Bins();
}

cudaError_t bins_build(Bins& b, Launcher& L, const CellGeom& cg, const PatchBin* d_patches, int n_patches,
                       const PatchBin* h_patches, const double* d_X, long long x_stride, const uint32_t* d_tie,
                       uint32_t tie_bound, int n_entries, int* d_cells_out, int* d_owner_out)
{
    int total_bricks = 0;
    for (int p = 0; p < n_patches; ++p) total_bricks = std::max(total_bricks, h_patches[p].brick_base + h_patches[p].nbricks);
    cudaError_t e = bins_reserve(b, n_entries, total_bricks);
    if (e != cudaSuccess) return e;
    b.n_entries = n_entries;
    b.total_bricks = total_bricks;
    const int cshift = 2 * cg.ndim;
    // tie ids (Lagrangian indices, 32-bit ints in the reference, LNodeIndex.h:187-189) are < tie_bound
    const int tie_bits = bits_for(tie_bound > 1 ? (unsigned long long)tie_bound - 1ull : 1ull);
    b.tie_bits = tie_bits;
    const int bkey_bits = bits_for(((unsigned long long)total_bricks << cshift) + 1ull);
    BinParams bp;
    bp.cg = cg;
    bp.n_patches = n_patches;
    bp.key_tie_bits = tie_bits;
    bp.cshift = cshift;
    const int T = 256;
    if (n_entries > 0)
    {
        bin_keys_kernel<<<(n_entries + T - 1) / T, T, 0, L.stream>>>(bp, d_patches, d_X, x_stride, d_tie, n_entries,
                                                                    b.keys[0], b.vals[0], d_cells_out, d_owner_out);
        L.launches++;
    }
    // sort by the brick / cell bits only (4 passes instead of 7 for 2^23 markers in 2.3 M bricks); the few markers that share
    // a cell are ordered by their tie bits afterwards
    const int end_bit = tie_bits + ((bkey_bits + 7) / 8) * 8;
    b.sorted_in = radix_sort_pairs(b.keys[0], b.vals[0], b.keys[1], b.vals[1], n_entries, tie_bits, end_bit, b.sort_temp,
                                   L.stream, &L.launches);
    if ((e = sort_ties(L, b.keys[b.sorted_in], b.vals[b.sorted_in], n_entries, tie_bits)) != cudaSuccess) return e;
    brick_offsets_kernel<<<(n_entries + 1 + T - 1) / T, T, 0, L.stream>>>(b.keys[b.sorted_in], n_entries, tie_bits + cshift,
                                                                         total_bricks, b.brick_start);
    L.launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // dense bricks: at most n / (thresh + 1) of them
    {
        const int cap = n_entries / (DENSE_BRICK_MARKERS + 1) + 1;
        if (cap > b.dense_capacity)
        {
            if (b.dense_list) cudaFree(b.dense_list);
            if ((e = cudaMalloc(&b.dense_list, sizeof(int) * (size_t)cap)) != cudaSuccess) return e;
            b.dense_capacity = cap;
        }
        if (!b.dense_count && (e = cudaMalloc(&b.dense_count, sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(b.dense_count, 0, sizeof(int), L.stream)) != cudaSuccess) return e;
        dense_bricks_kernel<<<(total_bricks + T - 1) / T, T, 0, L.stream>>>(b.brick_start, total_bricks, DENSE_BRICK_MARKERS, b.dense_list,
                                                                           b.dense_capacity, b.dense_count);
        L.launches++;
        if ((e = cudaMemcpyAsync(&b.n_dense, b.dense_count, sizeof(int), cudaMemcpyDeviceToHost, L.stream)) != cudaSuccess) return e;
    }
    // marker range of every patch (its bricks are one contiguous id range)
    b.range_base.assign(n_patches, 0);
    b.range_first.assign(n_patches, 0);
    b.range_last.assign(n_patches, 0);
    for (int p = 0; p < n_patches; ++p)
    {
        b.range_base[p] = h_patches[p].brick_base;
        if ((e = cudaMemcpyAsync(&b.range_first[p], b.brick_start + h_patches[p].brick_base, sizeof(int), cudaMemcpyDeviceToHost,
                                 L.stream)) != cudaSuccess)
            return e;
        if ((e = cudaMemcpyAsync(&b.range_last[p], b.brick_start + h_patches[p].brick_base + h_patches[p].nbricks, sizeof(int),
                                 cudaMemcpyDeviceToHost, L.stream)) != cudaSuccess)
            return e;
    }
    return cudaStreamSynchronize(L.stream);
}

cudaError_t wrap_positions(Launcher& L, const DomainGeom& dg, double* d_X, long long x_stride, int n, int* d_escaped, bool check_only)
{
    if (n <= 0) return cudaSuccess;
    const int T = 256;
    wrap_positions_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(dg, d_X, x_stride, n, d_escaped, check_only ? 1 : 0);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t gather_columns(Launcher& L, const double* d_in, long long in_stride, double* d_out, long long out_stride,
                           const uint32_t* d_perm, int n, int ncols)
{
    if (n <= 0) return cudaSuccess;
    const int T = 256;
    gather_columns_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(d_in, in_stride, d_out, out_stride, d_perm, n, ncols);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t gather_column_sets(Launcher& L, const GatherSets& sets, long long stride, const uint32_t* d_perm, int n, int ncols)
{
    if (n <= 0 || sets.nsets <= 0) return cudaSuccess;
    const int T = 256;
    gather_column_sets_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(sets, stride, d_perm, n, ncols);
    L.launches++;
    return cudaGetLastError();
}
cudaError_t sort_ties(Launcher& L, uint64_t* keys, uint32_t* vals, int n, int tie_bits)
{
    if (n <= 1) return cudaSuccess;
    const int T = 256;
    sort_ties_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(keys, vals, n, tie_bits);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t scatter_columns(Launcher& L, const double* d_in, long long in_stride, double* d_out, long long out_stride,
                            const uint32_t* d_perm, int n, int ncols)
{
    if (n <= 0) return cudaSuccess;
    const int T = 256;
    scatter_columns_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(d_in, in_stride, d_out, out_stride, d_perm, n, ncols);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t extract_low32(Launcher& L, const uint64_t* d_keys, uint32_t* d_out, int n, int bits)
{
    if (n <= 0) return cudaSuccess;
    const int T = 256;
    extract_low_kernel<<<(n + T - 1) / T, T, 0, L.stream>>>(d_keys, d_out, n, (1ull << bits) - 1ull);
    L.launches++;
    return cudaGetLastError();
}

} // namespace ibk
