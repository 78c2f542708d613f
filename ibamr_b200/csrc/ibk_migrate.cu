// ibk_migrate.cu -- marker migration between ranks at a re-distribution (device side).
//
// Role in the reference: LDataManager::endDataRedistribution moves every LData row of a node whose cell
// changed owner to the new owner's process through a PETSc VecScatter built from the new node distribution
// (LDataManager.cpp:1519-1959, scatter :1824-1837; computeNodeDistribution :2857-3046).  Here the owner-only
// binning (ibk_bin.cu) already parks the markers no local patch accepts at the tail of the sorted order; this
// file finds the destination rank of each of them from the level's global box list, groups them by
// destination (stable radix sort), packs [X, U, F, Lagrangian index] rows for an all-to-all, and appends the
// rows that arrive.  The all-to-all itself is the host's (torch.distributed / MPI): ibamr_b200/halo.py.
#include <cuda_runtime.h>

#include "ibk_device.cuh"
#include "ibk_engine.h"

namespace ibk
{
__global__ void migrate_dest_kernel(CellGeom cg, const int* __restrict__ plo, const int* __restrict__ phi,
                                    const int* __restrict__ prank, int n_patches, int n_ranks, const double* __restrict__ X,
                                    long long stride, int first, int n_tail, uint64_t* __restrict__ keys,
                                    uint32_t* __restrict__ vals)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_tail) return;
    const int i = first + j;
    int c[3] = { 0, 0, 0 };
    for (int d = 0; d < cg.ndim; ++d) // IndexUtilities::getCellIndex, as in bin_keys_kernel
        c[d] = cell_index_1d(X[d * stride + i], cg.x_lower[d], cg.x_upper[d], cg.dx[d], cg.ilower[d], cg.iupper[d]);
    int dest = n_ranks; // bucket "no patch of the level holds this cell"
    for (int p = 0; p < n_patches; ++p)
    {
        bool in = true;
        for (int d = 0; d < cg.ndim; ++d) in = in && c[d] >= plo[p * cg.ndim + d] && c[d] <= phi[p * cg.ndim + d];
        if (in)
        {
            dest = prank[p];
            break;
        }
    }
    keys[j] = (uint64_t)dest;
    vals[j] = (uint32_t)i;
}

// start[b] = first sorted position whose key is >= b, b = 0..n_buckets (start[n_buckets] = n)
__global__ void bucket_offsets_kernel(const uint64_t* __restrict__ keys, int n, int n_buckets, int* __restrict__ start)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long long prev = (i == 0) ? -1 : (long long)keys[i - 1];
    const long long cur = (i == n) ? n_buckets : (long long)keys[i];
    for (long long b = prev + 1; b <= cur; ++b) start[b] = i;
}

__global__ void migrate_pack_kernel(const uint32_t* __restrict__ order, int n_send, const double* __restrict__ X,
                                    const double* __restrict__ U, const double* __restrict__ F, long long stride, int ndim,
                                    const uint32_t* __restrict__ gid, double* __restrict__ buf)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_send) return;
    const int i = (int)order[j];
    double* row = buf + (size_t)j * (3 * ndim + 1);
    for (int d = 0; d < ndim; ++d)
    {
        row[d] = X[d * stride + i];
        row[ndim + d] = U[d * stride + i];
        row[2 * ndim + d] = F[d * stride + i];
    }
    row[3 * ndim] = (double)gid[i];
}

__global__ void migrate_append_kernel(const double* __restrict__ buf, int n_recv, int at, double* __restrict__ X,
                                      double* __restrict__ U, double* __restrict__ F, long long stride, int ndim,
                                      uint32_t* __restrict__ gid)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_recv) return;
    const double* row = buf + (size_t)j * (3 * ndim + 1);
    const int i = at + j;
    for (int d = 0; d < ndim; ++d)
    {
        X[d * stride + i] = row[d];
        U[d * stride + i] = row[ndim + d];
        F[d * stride + i] = row[2 * ndim + d];
    }
    gid[i] = (uint32_t)row[3 * ndim];
}

__global__ void id_keys_kernel(const uint32_t* __restrict__ gid, int n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = gid[i];
    vals[i] = (uint32_t)i;
}
__global__ void rank_scatter_kernel(const uint32_t* __restrict__ sorted_pos, int n, uint32_t* __restrict__ row)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    row[sorted_pos[k]] = (uint32_t)k;
}
__global__ void gather_u32_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, int n, uint32_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[perm[i]];
}

static inline unsigned nblk(int n)
{
    return (unsigned)((n + 255) / 256);
}

cudaError_t migrate_dest(Launcher& L, const CellGeom& cg, const int* d_plo, const int* d_phi, const int* d_prank, int n_patches,
                         int n_ranks, const double* X, long long stride, int first, int n_tail, uint64_t* keys, uint32_t* vals)
{
    if (n_tail <= 0) return cudaSuccess;
    migrate_dest_kernel<<<nblk(n_tail), 256, 0, L.stream>>>(cg, d_plo, d_phi, d_prank, n_patches, n_ranks, X, stride, first,
                                                            n_tail, keys, vals);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t bucket_offsets(Launcher& L, const uint64_t* keys_sorted, int n, int n_buckets, int* d_start)
{
    bucket_offsets_kernel<<<nblk(n + 1), 256, 0, L.stream>>>(keys_sorted, n, n_buckets, d_start);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t migrate_pack(Launcher& L, const uint32_t* order, int n_send, const double* X, const double* U, const double* F,
                         long long stride, int ndim, const uint32_t* gid, double* buf)
{
    if (n_send <= 0) return cudaSuccess;
    migrate_pack_kernel<<<nblk(n_send), 256, 0, L.stream>>>(order, n_send, X, U, F, stride, ndim, gid, buf);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t migrate_append(Launcher& L, const double* buf, int n_recv, int at, double* X, double* U, double* F, long long stride,
                           int ndim, uint32_t* gid)
{
    if (n_recv <= 0) return cudaSuccess;
    migrate_append_kernel<<<nblk(n_recv), 256, 0, L.stream>>>(buf, n_recv, at, X, U, F, stride, ndim, gid);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t id_keys(Launcher& L, const uint32_t* gid, int n, uint64_t* keys, uint32_t* vals)
{
    if (n <= 0) return cudaSuccess;
    id_keys_kernel<<<nblk(n), 256, 0, L.stream>>>(gid, n, keys, vals);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t rank_scatter(Launcher& L, const uint32_t* sorted_pos, int n, uint32_t* row)
{
    if (n <= 0) return cudaSuccess;
    rank_scatter_kernel<<<nblk(n), 256, 0, L.stream>>>(sorted_pos, n, row);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t gather_u32(Launcher& L, const uint32_t* in, const uint32_t* perm, int n, uint32_t* out)
{
    if (n <= 0) return cudaSuccess;
    gather_u32_kernel<<<nblk(n), 256, 0, L.stream>>>(in, perm, n, out);
    L.launches++;
    return cudaGetLastError();
}

} // namespace ibk
