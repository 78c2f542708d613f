// ibk_sort.cu -- stable LSD radix sort of (uint64 key, uint32 value) pairs, written for this
// library (no CUB/Thrust).  It is the device replacement for the ordering work of
// LDataManager::beginDataRedistribution / computeNodeDistribution
// (ibtk/src/lagrangian/LDataManager.cpp:1465-1508, 2897-2911), where the reference uses
// std::sort per cell + std::map per marker.
//
// 8-bit digits, three kernels per pass:
//   radix_hist    one CTA per 4096-key chunk: 256-bin histogram -> hist[bin][chunk]
//   radix_scan    one CTA per bin: exclusive scan over chunks; radix_scan_bins: scan of the 256 totals
//   radix_scatter one CTA per chunk, each warp owns a contiguous 512-key slice and ranks its keys
//                 32 at a time with __match_any_sync, so equal digits keep their input order
// Passes whose digit is identical for every key (one bin holds all n keys) are skipped; the host
// never reads anything back: the skip decision is taken on the device (flag in `ctrl`), the
// scatter of a skipped pass degenerates to a straight copy so the ping-pong parity stays fixed.
// All traffic is HBM-bound integer work: 8+4 bytes read twice and written once per key per pass.
#include <cuda_runtime.h>
#include <stdint.h>

namespace ibk
{
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS_PER_WARP = 512;
constexpr int RS_CHUNK = RS_WARPS * RS_ITEMS_PER_WARP; // 4096 keys per CTA
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_THREADS)
    radix_hist(const uint64_t* __restrict__ keys, int n, int shift, unsigned* __restrict__ hist, int nchunks)
{
    __shared__ unsigned sh[RS_BINS];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_CHUNK;
    const int end = min(base + RS_CHUNK, n);
    for (int i = base + threadIdx.x; i < end; i += RS_THREADS)
    {
        const unsigned d = (unsigned)(keys[i] >> shift) & 0xFFu;
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nchunks + blockIdx.x] = sh[threadIdx.x];
}

// One CTA per bin: exclusive scan of hist[bin][0..nchunks) in place, total -> bin_total[bin].
__global__ void __launch_bounds__(RS_THREADS) radix_scan(unsigned* __restrict__ hist, int nchunks, unsigned* __restrict__ bin_total)
{
    __shared__ unsigned warp_sums[RS_WARPS];
    __shared__ unsigned carry;
    unsigned* row = hist + (size_t)blockIdx.x * nchunks;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nchunks; base += RS_THREADS)
    {
        const int i = base + threadIdx.x;
        const unsigned v = (i < nchunks) ? row[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        unsigned woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
            if (w < warp) woff += warp_sums[w];
        const unsigned c = carry;
        if (i < nchunks) row[i] = c + woff + x - v;
        __syncthreads();
        if (threadIdx.x == RS_THREADS - 1) carry = c + woff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = carry;
}

// Exclusive scan of the 256 bin totals; ctrl[0] = 1 iff one bin holds every key (pass is a no-op).
__global__ void __launch_bounds__(RS_BINS) radix_scan_bins(unsigned* __restrict__ bin_total, int n, int* __restrict__ ctrl)
{
    __shared__ unsigned s[RS_BINS];
    __shared__ int skip;
    const unsigned v = bin_total[threadIdx.x];
    s[threadIdx.x] = v;
    if (threadIdx.x == 0) skip = 0;
    __syncthreads();
    if (v == (unsigned)n && n > 0) skip = 1;
    // Hillis-Steele inclusive scan
    for (int o = 1; o < RS_BINS; o <<= 1)
    {
        unsigned y = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += y;
        __syncthreads();
    }
    bin_total[threadIdx.x] = s[threadIdx.x] - v;
    if (threadIdx.x == 0) ctrl[0] = skip;
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out,
                                                            int n,
                                                            int shift,
                                                            const unsigned* __restrict__ hist,
                                                            const unsigned* __restrict__ bin_base,
                                                            const int* __restrict__ ctrl,
                                                            int nchunks)
{
    __shared__ unsigned cnt[RS_WARPS][RS_BINS]; // per-warp digit counts, then running offsets
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk_base = blockIdx.x * RS_CHUNK;
    if (ctrl[0])
    {
        // every key has the same digit: identity permutation, keep the ping-pong going
        const int end = min(chunk_base + RS_CHUNK, n);
        for (int i = chunk_base + threadIdx.x; i < end; i += RS_THREADS)
        {
            keys_out[i] = keys_in[i];
            vals_out[i] = vals_in[i];
        }
        return;
    }
    for (int b = lane; b < RS_BINS; b += 32) cnt[warp][b] = 0;
    __syncwarp();
    const int wbase = chunk_base + warp * RS_ITEMS_PER_WARP;
    const int wend = min(wbase + RS_ITEMS_PER_WARP, n);
    // pass 1: this warp's digit histogram
    for (int i = wbase + lane; i < wend; i += 32)
    {
        const unsigned d = (unsigned)(keys_in[i] >> shift) & 0xFFu;
        atomicAdd(&cnt[warp][d], 1u);
    }
    __syncthreads();
    // per digit: exclusive prefix over warps + global bases -> running offsets
    {
        const int b = threadIdx.x; // RS_THREADS == RS_BINS
        unsigned run = bin_base[b] + hist[(size_t)b * nchunks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
        {
            const unsigned c = cnt[w][b];
            cnt[w][b] = run;
            run += c;
        }
    }
    __syncthreads();
    // pass 2: rank 32 keys at a time, in input order
    for (int i0 = wbase; i0 < wend; i0 += 32)
    {
        const int i = i0 + lane;
        const bool active = i < wend;
        uint64_t k = 0;
        uint32_t v = 0;
        unsigned d = 0xFFFFFFFFu; // inactive lanes get a digit no active lane has
        if (active)
        {
            k = keys_in[i];
            v = vals_in[i];
            d = (unsigned)(k >> shift) & 0xFFu;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned lower = peers & ((1u << lane) - 1u);
        if (active)
        {
            const unsigned pos = cnt[warp][d] + __popc(lower);
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
        __syncwarp();
        if (active && lower == 0) cnt[warp][d] += __popc(peers); // group leader advances the offset
        __syncwarp();
    }
}

size_t radix_sort_temp_bytes(int n)
{
    const int nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
    return sizeof(unsigned) * ((size_t)RS_BINS * (size_t)(nchunks > 0 ? nchunks : 1) + RS_BINS) + 64;
}

// Sorts by bits [begin_bit, end_bit) (multiples of 8).  Input in (keys_a, vals_a); (keys_b, vals_b)
// is scratch.  Returns 0 if the sorted result is in the *_a buffers, 1 if it is in *_b.
// Adds the number of kernel launches to *launches.
int radix_sort_pairs(uint64_t* keys_a,
                     uint32_t* vals_a,
                     uint64_t* keys_b,
                     uint32_t* vals_b,
                     int n,
                     int begin_bit,
                     int end_bit,
                     void* temp,
                     cudaStream_t stream,
                     long long* launches)
{
    if (n <= 0) return 0;
    const int nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
    unsigned* hist = (unsigned*)temp;
    unsigned* bin_total = hist + (size_t)RS_BINS * nchunks;
    int* ctrl = (int*)(bin_total + RS_BINS);
    int which = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8)
    {
        uint64_t* kin = which ? keys_b : keys_a;
        uint64_t* kout = which ? keys_a : keys_b;
        uint32_t* vin = which ? vals_b : vals_a;
        uint32_t* vout = which ? vals_a : vals_b;
        radix_hist<<<nchunks, RS_THREADS, 0, stream>>>(kin, n, shift, hist, nchunks);
        radix_scan<<<RS_BINS, RS_THREADS, 0, stream>>>(hist, nchunks, bin_total);
        radix_scan_bins<<<1, RS_BINS, 0, stream>>>(bin_total, n, ctrl);
        radix_scatter<<<nchunks, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, hist, bin_total, ctrl, nchunks);
        if (launches) *launches += 4;
        which ^= 1;
    }
    return which;
}

} // namespace ibk
