// ibk_halo.cu -- grid halo operations and layout conversions.
//
// Device replacements for the SAMRAI schedules around the hot path:
//   region copy        u ghost fill, u_ghost_fill_scheds[ln]->fillData (LDataManager.cpp:744):
//                      same-level copies incl. periodic images
//   region add         ghost-region accumulation onto the owning DOFs,
//                      SAMRAIGhostDataAccumulator::accumulateGhostData
//                      (ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:327-334, ADD_VALUES/SCATTER_REVERSE)
//   pack / unpack      the same regions through a contiguous buffer, for the inter-process exchange
// plus AoS<->SoA (LData seam) and dense<->pitched (SAMRAI ArrayData seam) conversions.
// All of it is pure HBM streaming: one thread per element, x fastest so that warps touch
// contiguous 256-byte runs.
#include <cuda_runtime.h>

#include <algorithm>

#include "ibk_engine.h"
#include "ibk_device.cuh"

namespace ibk
{
struct RegionBatch
{
    RegionCopy op[16];
    int n;
};

template <int MODE>
__global__ void region_ops_kernel(RegionBatch rb)
{
    const RegionCopy& r = rb.op[blockIdx.y];
    const long long total = (long long)r.ext[0] * r.ext[1] * r.ext[2];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q % r.ext[0]);
        const long long t = q / r.ext[0];
        const int j = (int)(t % r.ext[1]);
        const int k = (int)(t / r.ext[1]);
        const double v = r.src[((long long)(r.src_off[2] + k) * r.src_n1 + (r.src_off[1] + j)) * r.src_pitch + r.src_off[0] + i];
        double* p = r.dst + ((long long)(r.dst_off[2] + k) * r.dst_n1 + (r.dst_off[1] + j)) * r.dst_pitch + r.dst_off[0] + i;
        if (MODE == 0)
            *p = v;
        else
            *p += v;
    }
}

cudaError_t launch_region_ops(Launcher& L, const std::vector<RegionCopy>& ops, int mode)
{
    // The caller orders `ops` so that no two ops of one call write the same element (adds into the
    // same owner from different sources go into separate calls), which keeps the sums deterministic.
    for (size_t base = 0; base < ops.size(); base += 16)
    {
        RegionBatch rb;
        rb.n = (int)std::min<size_t>(16, ops.size() - base);
        long long maxtotal = 1;
        for (int i = 0; i < rb.n; ++i)
        {
            rb.op[i] = ops[base + i];
            const long long t = (long long)rb.op[i].ext[0] * rb.op[i].ext[1] * rb.op[i].ext[2];
            if (t > maxtotal) maxtotal = t;
        }
        const int T = 256;
        long long nbx = (maxtotal + T - 1) / T;
        if (nbx > 4096) nbx = 4096;
        dim3 grid((unsigned)nbx, (unsigned)rb.n);
        if (mode == 0)
            region_ops_kernel<0><<<grid, T, 0, L.stream>>>(rb);
        else
            region_ops_kernel<1><<<grid, T, 0, L.stream>>>(rb);
        L.launches++;
    }
    return cudaGetLastError();
}

__global__ void pack_kernel(const double* __restrict__ arr, long long pitch, int n1, int o0, int o1, int o2, int e0, int e1,
                            int e2, double* __restrict__ buf)
{
    const long long total = (long long)e0 * e1 * e2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q % e0);
        const long long t = q / e0;
        const int j = (int)(t % e1);
        const int k = (int)(t / e1);
        buf[q] = arr[((long long)(o2 + k) * n1 + (o1 + j)) * pitch + o0 + i];
    }
}

template <int MODE>
__global__ void unpack_kernel(double* __restrict__ arr, long long pitch, int n1, int o0, int o1, int o2, int e0, int e1,
                              int e2, const double* __restrict__ buf)
{
    const long long total = (long long)e0 * e1 * e2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q % e0);
        const long long t = q / e0;
        const int j = (int)(t % e1);
        const int k = (int)(t / e1);
        double* p = arr + ((long long)(o2 + k) * n1 + (o1 + j)) * pitch + o0 + i;
        if (MODE == 0)
            *p = buf[q];
        else
            *p += buf[q];
    }
}

static unsigned grid_for(long long total, int T)
{
    long long nb = (total + T - 1) / T;
    if (nb > 148 * 32) nb = 148 * 32;
    if (nb < 1) nb = 1;
    return (unsigned)nb;
}

cudaError_t launch_pack(Launcher& L, const double* arr, long long pitch, int n1, const int* off, const int* ext, double* buf,
                        int ndim)
{
    const int o2 = ndim == 3 ? off[2] : 0, e2 = ndim == 3 ? ext[2] : 1;
    const long long total = (long long)ext[0] * ext[1] * e2;
    if (total <= 0) return cudaSuccess;
    pack_kernel<<<grid_for(total, 256), 256, 0, L.stream>>>(arr, pitch, n1, off[0], off[1], o2, ext[0], ext[1], e2, buf);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t launch_unpack(Launcher& L, double* arr, long long pitch, int n1, const int* off, const int* ext, const double* buf,
                          int ndim, int mode)
{
    const int o2 = ndim == 3 ? off[2] : 0, e2 = ndim == 3 ? ext[2] : 1;
    const long long total = (long long)ext[0] * ext[1] * e2;
    if (total <= 0) return cudaSuccess;
    if (mode == 0)
        unpack_kernel<0><<<grid_for(total, 256), 256, 0, L.stream>>>(arr, pitch, n1, off[0], off[1], o2, ext[0], ext[1], e2, buf);
    else
        unpack_kernel<1><<<grid_for(total, 256), 256, 0, L.stream>>>(arr, pitch, n1, off[0], off[1], o2, ext[0], ext[1], e2, buf);
    L.launches++;
    return cudaGetLastError();
}

long long region_work(const int* ext)
{
    const long long rows = (long long)ext[1] * ext[2];
    if (ext[0] >= 16) return rows * ((ext[0] + 127) / 128);
    const int rpw = 32 / std::max(ext[0], 1);
    return (rows + rpw - 1) / rpw;
}
unsigned region_blocks(const int* ext)
{
    return (unsigned)std::max<long long>(1, (region_work(ext) + 15) / 16); // 16 warp work items per CTA of 8 warps
}

// All regions of a message (or of one wave of it) in one launch: the CTAs are dealt out to the items in proportion to
// their size, a warp walks a row of its region (contiguous in the array and in the buffer) or a few short rows.
__global__ void halo_items_kernel(const HaloItem* __restrict__ items, int n_items, double* __restrict__ buf, int op)
{
    int lo_i = 0, hi_i = n_items - 1; // the last item with block0 <= blockIdx.x
    while (lo_i < hi_i)
    {
        const int mid = (lo_i + hi_i + 1) >> 1;
        if (items[mid].block0 <= blockIdx.x) lo_i = mid;
        else hi_i = mid - 1;
    }
    const HaloItem& it = items[lo_i];
    const int lo[3] = { 0, 0, 0 }, hi[3] = { it.ext[0] - 1, it.ext[1] - 1, it.ext[2] - 1 };
    double* b = buf + it.buf_off;
    slab_rows(lo, hi, blockIdx.x - it.block0, it.nblocks, [&](int j, int k, int x0, int x1) {
        double* arow = it.ptr + ((long long)(it.off[2] + k) * it.n1 + (it.off[1] + j)) * it.pitch + it.off[0];
        double* brow = b + ((long long)k * it.ext[1] + j) * it.ext[0];
        for (int x = x0; x <= x1; x += 32)
        {
            if (op == 0) brow[x] = arow[x];
            else if (op == 1) arow[x] = brow[x];
            else arow[x] += brow[x];
        }
    });
}

cudaError_t launch_halo_items(Launcher& L, const HaloItem* d_items, int n_items, unsigned total_blocks, double* buf, int op)
{
    if (n_items <= 0 || total_blocks == 0) return cudaSuccess;
    halo_items_kernel<<<total_blocks, 256, 0, L.stream>>>(d_items, n_items, buf, op);
    L.launches++;
    return cudaGetLastError();
}

__global__ void fill_kernel(double* __restrict__ p, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

cudaError_t launch_fill(Launcher& L, double* ptr, size_t count, double value)
{
    if (count == 0) return cudaSuccess;
    if (value == 0.0) return cudaMemsetAsync(ptr, 0, count * sizeof(double), L.stream);
    fill_kernel<<<grid_for((long long)count, 256), 256, 0, L.stream>>>(ptr, count, value);
    L.launches++;
    return cudaGetLastError();
}

// Host <-> pitched device array.  cudaMemcpy2DAsync moves host->device rows of a few KB at about half the
// link rate (measured: 30 GB/s against 55 GB/s for a flat copy of the same bytes; scripts/pcie_probe.py), so
// with a staging buffer the bytes cross the link as flat copies of dense row blocks and a small kernel moves
// them between the dense block and the pitched array (on the same stream: in order, no extra events).
__global__ void repitch_kernel(const double* __restrict__ src, double* __restrict__ dst, int n0, long long src_pitch,
                               long long dst_pitch, long long rows)
{
    const long long row = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    if (row >= rows) return;
    const double* s = src + row * src_pitch;
    double* d = dst + row * dst_pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) d[i] = s[i];
}

static cudaError_t launch_repitch(Launcher& L, const double* src, double* dst, int n0, long long src_pitch, long long dst_pitch,
                                  long long rows)
{
    const dim3 block(128, 2);
    const long long by = (rows + block.y - 1) / block.y;
    for (long long r0 = 0; r0 < by; r0 += 65535) // gridDim.y limit
    {
        const long long nby = std::min<long long>(65535, by - r0);
        const long long row0 = r0 * block.y;
        repitch_kernel<<<dim3((unsigned)std::max(1, std::min(8, (n0 + 127) / 128)), (unsigned)nby), block, 0, L.stream>>>(
            src + row0 * src_pitch, dst + row0 * dst_pitch, n0, src_pitch, dst_pitch, std::min<long long>(rows - row0, nby * block.y));
        L.launches++;
    }
    return cudaGetLastError();
}

cudaError_t copy_dense_to_pitched(Launcher& L, const double* src_dense, double* dst, long long pitch, const int* n, int ndim,
                                  cudaMemcpyKind kind, void* stage, size_t stage_bytes)
{
    const size_t rows = (size_t)n[1] * (ndim == 3 ? n[2] : 1);
    const size_t row_bytes = (size_t)n[0] * sizeof(double);
    if (pitch == n[0]) return cudaMemcpyAsync(dst, src_dense, rows * row_bytes, kind, L.stream);
    if (!stage || stage_bytes < row_bytes)
        return cudaMemcpy2DAsync(dst, (size_t)pitch * sizeof(double), src_dense, row_bytes, row_bytes, rows, kind, L.stream);
    const size_t rows_per_chunk = stage_bytes / row_bytes;
    cudaError_t e;
    for (size_t r = 0; r < rows; r += rows_per_chunk)
    {
        const size_t nr = std::min(rows_per_chunk, rows - r);
        if ((e = cudaMemcpyAsync(stage, src_dense + r * n[0], nr * row_bytes, kind, L.stream)) != cudaSuccess) return e;
        if ((e = launch_repitch(L, (const double*)stage, dst + r * pitch, n[0], n[0], pitch, (long long)nr)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t copy_pitched_to_dense(Launcher& L, const double* src, long long pitch, double* dst_dense, const int* n, int ndim,
                                  cudaMemcpyKind kind, void* stage, size_t stage_bytes)
{
    const size_t rows = (size_t)n[1] * (ndim == 3 ? n[2] : 1);
    const size_t row_bytes = (size_t)n[0] * sizeof(double);
    if (pitch == n[0]) return cudaMemcpyAsync(dst_dense, src, rows * row_bytes, kind, L.stream);
    if (!stage || stage_bytes < row_bytes)
        return cudaMemcpy2DAsync(dst_dense, row_bytes, src, (size_t)pitch * sizeof(double), row_bytes, rows, kind, L.stream);
    const size_t rows_per_chunk = stage_bytes / row_bytes;
    cudaError_t e;
    for (size_t r = 0; r < rows; r += rows_per_chunk)
    {
        const size_t nr = std::min(rows_per_chunk, rows - r);
        if ((e = launch_repitch(L, src + r * pitch, (double*)stage, n[0], pitch, n[0], (long long)nr)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(dst_dense + r * n[0], stage, nr * row_bytes, kind, L.stream)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, long long stride, int n, int depth)
{
    const long long total = (long long)n * depth;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q / depth), d = (int)(q % depth);
        soa[d * stride + i] = aos[q];
    }
}
__global__ void soa_to_aos_kernel(const double* __restrict__ soa, long long stride, double* __restrict__ aos, int n, int depth)
{
    const long long total = (long long)n * depth;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(q / depth), d = (int)(q % depth);
        aos[q] = soa[d * stride + i];
    }
}

cudaError_t aos_to_soa(Launcher& L, const double* d_aos, double* d_soa, long long stride, int n, int depth)
{
    if (n <= 0) return cudaSuccess;
    aos_to_soa_kernel<<<grid_for((long long)n * depth, 256), 256, 0, L.stream>>>(d_aos, d_soa, stride, n, depth);
    L.launches++;
    return cudaGetLastError();
}
cudaError_t soa_to_aos(Launcher& L, const double* d_soa, long long stride, double* d_aos, int n, int depth)
{
    if (n <= 0) return cudaSuccess;
    soa_to_aos_kernel<<<grid_for((long long)n * depth, 256), 256, 0, L.stream>>>(d_soa, stride, d_aos, n, depth);
    L.launches++;
    return cudaGetLastError();
}

__global__ void build_entries_kernel(const double* __restrict__ X, const int* __restrict__ idx, const double* __restrict__ shift,
                                     int n, int ndim, double* __restrict__ Xe, double* __restrict__ Xr, long long stride)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const long long s = idx ? idx[l] : l;
    for (int d = 0; d < ndim; ++d)
    {
        const double x = X[s * ndim + d];
        const double sh = shift ? shift[(long long)l * ndim + d] : 0.0;
        Xe[d * stride + l] = __dadd_rn(x, sh); // X(d,s)+Xshift(d,l), 3d.f.m4:1265
        if (Xr) Xr[d * stride + l] = x;
    }
}

cudaError_t build_entries(Launcher& L, const double* d_X_aos, const int* d_idx, const double* d_shift, int n, int ndim,
                          double* d_Xe, double* d_Xr, long long stride)
{
    if (n <= 0) return cudaSuccess;
    build_entries_kernel<<<(n + 255) / 256, 256, 0, L.stream>>>(d_X_aos, d_idx, d_shift, n, ndim, d_Xe, d_Xr, stride);
    L.launches++;
    return cudaGetLastError();
}

// Listed markers that no tile can reach (binned into the discard bucket) get V = 0, as the Fortran
// does for a fully clipped stencil (V(d,s) = 0.d0 before the empty loops, 3d.f.m4:1316).
__global__ void zero_discarded_kernel(const int* __restrict__ brick_start, int total_bricks, int n, const uint32_t* __restrict__ src,
                                      double* __restrict__ V, long long v_cstride, long long v_istride, int ncol)
{
    const int first = brick_start[total_bricks];
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const long long row = src ? (long long)src[i] : (long long)i;
        for (int c = 0; c < ncol; ++c) V[c * v_cstride + row * v_istride] = 0.0;
    }
}

cudaError_t zero_discarded(Launcher& L, const int* brick_start, int total_bricks, int n, const uint32_t* src, double* V,
                           long long v_cstride, long long v_istride, int ncol)
{
    if (n <= 0) return cudaSuccess;
    zero_discarded_kernel<<<64, 256, 0, L.stream>>>(brick_start, total_bricks, n, src, V, v_cstride, v_istride, ncol);
    L.launches++;
    return cudaGetLastError();
}

// Position-only interpolation: a marker whose cell lies in the caller's box but outside the range the binning accepts
// (farther than gcw + 4 cells from the patch) has no array point under its stencil: the reference lists it
// (LEInteractor::buildLocalIndices), interpolates a fully clipped stencil and writes 0 (LEInteractor.cpp:3117-3120).
struct BoxTest
{
    CellGeom cg;
    int lo[3], hi[3];
};
__global__ void zero_discarded_in_box_kernel(const int* __restrict__ brick_start, int total_bricks, int n, const double* __restrict__ Xs,
                                             long long stride, BoxTest bt, const uint32_t* __restrict__ src, double* __restrict__ V,
                                             long long v_cstride, long long v_istride, int ncol)
{
    const int first = brick_start[total_bricks];
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        bool in = true;
        for (int d = 0; d < bt.cg.ndim; ++d)
        {
            const double x = Xs[d * stride + i];
            const int c = bt.cg.two_branch ? cell_index_1d(x, bt.cg.x_lower[d], bt.cg.x_upper[d], bt.cg.dx[d], bt.cg.ilower[d], bt.cg.iupper[d]) :
                                             bt.cg.ilower[d] + (int)floor(__ddiv_rn(__dsub_rn(x, bt.cg.x_lower[d]), bt.cg.dx[d]));
            in = in && c >= bt.lo[d] && c <= bt.hi[d];
        }
        if (!in) continue;
        const long long row = src ? (long long)src[i] : (long long)i;
        for (int c = 0; c < ncol; ++c) V[c * v_cstride + row * v_istride] = 0.0;
    }
}
cudaError_t zero_discarded_in_box(Launcher& L, const int* brick_start, int total_bricks, int n, const double* Xs, long long stride,
                                  const CellGeom& cg, const int* box_lo, const int* box_hi, const uint32_t* src, double* V,
                                  long long v_cstride, long long v_istride, int ncol)
{
    if (n <= 0) return cudaSuccess;
    BoxTest bt;
    bt.cg = cg;
    for (int d = 0; d < 3; ++d)
    {
        bt.lo[d] = d < cg.ndim ? box_lo[d] : 0;
        bt.hi[d] = d < cg.ndim ? box_hi[d] : 0;
    }
    zero_discarded_in_box_kernel<<<64, 256, 0, L.stream>>>(brick_start, total_bricks, n, Xs, stride, bt, src, V, v_cstride, v_istride, ncol);
    L.launches++;
    return cudaGetLastError();
}

__global__ void compose_index_kernel(const int* __restrict__ idx, const uint32_t* __restrict__ perm, uint32_t* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx ? (uint32_t)idx[perm[i]] : perm[i];
}

cudaError_t compose_index(Launcher& L, const int* d_idx, const uint32_t* d_perm, uint32_t* d_out, int n)
{
    if (n <= 0) return cudaSuccess;
    compose_index_kernel<<<(n + 255) / 256, 256, 0, L.stream>>>(d_idx, d_perm, d_out, n);
    L.launches++;
    return cudaGetLastError();
}

} // namespace ibk
