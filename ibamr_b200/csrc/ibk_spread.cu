// ibk_spread.cu -- force spreading (markers -> grid) for sm_100a, deterministic.
//
// Replaces lagrangian_<kernel>_spread{2,3}d (ibtk/src/lagrangian/fortran/
// lagrangian_interaction3d.f.m4:1344-1475 ib_4, :2384-2582 ib_6, :2703-2805 bspline_3,
// :2933-3041 bspline_4, :617-730 piecewise_linear; 2D twins in lagrangian_interaction2d.f.m4) and
// the per-axis loop of LEInteractor (LEInteractor.cpp:3676-3711).  The reference is serial over
// markers, so it has no write conflicts; here the work is organised so that none can occur and the
// order of the additions at every grid point is fixed (bit-reproducible, no float atomics):
//
//  3D, kernels with a reach of at most 3 cells (all but IB_4_W8): spread_march_kernel.  One CTA owns a COLUMN
//    of 32 x 32 cells and a chunk of marker tiles in z, and marches through it one brick layer (4 cells in z) at
//    a time.  The accumulator is a RING of z planes of (32 + 2M)^2 points in shared memory; when a layer is done
//    its four lowest planes are final and are ADDED to f by TMA's reducing store (cp.reduce.async.bulk.tensor
//    .add: no load of f, the addition happens in L2) while the next layer is accumulated, then zeroed and reused.
//    The CTA is warp-specialised: producer warps evaluate the 1-D stencils of layer s + 1 (one thread per marker,
//    three independent div/sqrt chains), consumer warps accumulate layer s, one warp flushes the planes of
//    layer s - 1.  A consumer warp takes one ROW of eight bricks along x and walks the bricks NC apart
//    (disjoint footprints) as independent chains side by side; rows NC apart are disjoint, so a layer takes NC
//    consumer-only barriers.  March tiles 2 apart are disjoint: 8 launches (colours).  f moves
//    (36/32 in y) x (40/32 in x, 32-byte sectors) x (68/64 in z) = 1.5 times instead of 2.3 times with 16^3 tiles.
//  2D and IB_4_W8: spread_tile_kernel.  One CTA per (marker tile of 16^ndim cells, component) with a haloed
//    block in shared memory loaded from / stored to f by TMA, brick colours inside, 2^ndim tile colours.
//  Dense bricks (3D, > 48 markers): spread_dense_kernel, register footprints.
//
// An (entry, component) pair whose stencil does not fit the accumulator of its tile (its binning cell and its
// stencil origin disagree: positions moved since the binning, or a rounding) is flagged in Bins::exc_flags and
// spread by spread_fixup_kernel in sorted order after the last colour: exact, never dropped, slow.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ibk_engine.h"
#include "ibk_tma.h"

namespace ibk
{
constexpr int SPREAD_THREADS = 256;

struct SpreadArgs
{
    const int* brick_start;
    const double* X;
    const double* Xraw;
    long long x_stride;
    const double* V;
    long long v_cstride, v_istride;
    const uint32_t* src;
    int colour[3]; // tile colour (parity per dimension) handled by this launch
    int ntc[3];    // number of tiles of that colour per dimension
    int* exc_count;
    unsigned* exc_flags;
    int n_entries;
    int cap; // markers whose stencil weights are staged at a time (sizes the dynamic shared memory)
    unsigned tma_mask; // bit a: the block of component a is moved by TMA
    int part, sel_lo[3], sel_hi[3]; // MarkerView's tile selection
    int clip_free;                  // MarkerView::clip_free
    int dense_thresh;               // > 0: bricks with more markers than this are left to spread_dense_kernel
    int chunk_tiles;                // march kernel: marker tiles per chunk in z
    // march kernel, persistent CTAs: work items = (march tile, component) in colour-major order, taken by ticket
    int nm[3];          // march tiles per dimension
    int item_base[9];   // first item of each tile colour (colour = x parity + 2 y parity + 4 z parity)
    int* work_counter;  // next ticket
    int* done_flags;    // [march tile][component]: the item has been added to f completely
};

__device__ __forceinline__ void flag_exception(const SpreadArgs& args, int i, int a)
{
    atomicOr(&args.exc_flags[i >> 2], 1u << (8 * (i & 3) + a));
    atomicAdd(args.exc_count, 1);
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// shared-memory accesses by 32-bit shared-space address (no generic-address arithmetic in the inner loop)
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_v2s32(uint32_t a)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// =============================================================================================
// 3D march kernel
// =============================================================================================
// a plane of zeros: the flusher refills flushed planes from it by bulk copies instead of storing zeros itself
__device__ __align__(128) double g_march_zeros[2048];
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#ifndef IBK_MARCH_NBAR
#define IBK_MARCH_NBAR 1
#endif
#ifndef IBK_MARCH_SUBARRIVE
#define IBK_MARCH_SUBARRIVE 1
#endif
constexpr int MARCH_CELLS = 2 * TILE;          // column edge: 2 x 2 marker tiles
constexpr int MARCH_BR = MARCH_CELLS / BRICK;  // 8 bricks per row / rows per layer
constexpr int MARCH_MAX_LAYERS = 32;

template <int K>
struct MarchCfg
{
    static constexpr int W = KTraits<K>::W;
    static constexpr int M = KTraits<K>::M;
    static constexpr int R = MARCH_CELLS + 2 * M; // haloed plane edge (y, and x before alignment)
    static constexpr int XO = M & 1;              // TMA boxes of 8-byte elements start on an even x coordinate
    static constexpr int RX = (R + XO + 1) & ~1;
    static constexpr int PLANE = (R * RX + 15) & ~15; // points per plane slot (a TMA source must start on a 128-byte boundary)
    static constexpr int FZ = BRICK + 2 * M;      // planes a brick layer reaches
    static constexpr int NRING = FZ + BRICK;      // + the four planes being flushed
    static constexpr int NC = (BRICK + 2 * M + BRICK - 1) / BRICK; // brick colours per dimension
    static constexpr int NCW = (MARCH_BR + NC - 1) / NC;           // rows per row colour = chains per row and sub-phase
    static constexpr bool FAST4 = (W == 4); // 64 points = 2 slots of 32 lanes with the same (x, y) per lane
    // FAST4: a row is shared by a PAIR of consumer warps (two of the four chains each): 8 consumer warps, two per SM
    // sub-partition, so one warp's shared-memory latency is covered by the other's issue slots
    static constexpr int WPR = FAST4 ? 2 : 1;                      // consumer warps per row
    static constexpr int NCONS = NCW * WPR;                        // consumer warps
    static constexpr int CPW = NCW / WPR;                          // chains per consumer warp
    static constexpr int WARPS = FAST4 ? 20 : 16;                  // 5 / 4 warps per SM sub-partition: 96 / 128 registers
    static constexpr int NPW = WARPS - NCONS - 1;                  // producer warps
    static constexpr int NT = 32 * WARPS;
    static constexpr int NPTS = W * W * W;
    static constexpr int NSLOT = (NPTS + 31) / 32;
    // per-marker record: W byte offsets (one per z plane of the stencil; negative: skip), then 3 x W weights
    static constexpr int ZO_BYTES = ((W * 4 + 15) / 16) * 16;
    // FAST4 (W = 4): 128 bytes.  Lane group g (= lane / 16, stencil planes g and g + 2) reads one 32-byte block at 32 g:
    // {int offset(g), int offset(g + 2), 8 spare bytes, double wz(g), double wz(g + 2)}; wx[4] at 64, wy[4] at 96.
    // (+ 16 spare bytes: with a stride of 128 the producers' stores of 32 records would all hit the same banks)
    static constexpr int REC = (W == 4) ? 144 : ((ZO_BYTES + 3 * W * 8 + 15) / 16) * 16;
    static constexpr int SINK_B = ((8 * (3 * RX + 4) + 127) / 128) * 128; // where the lanes of a dummy record add their zeros
};

// FAST4 consumer body: iterations [k, kend) of N chains side by side.  adr[c]: this lane's 32-byte block of chain c's
// current record; ring_lane: the ring plus the lane's (x, y) offset.  All record loads, then all ring loads, then all
// ring stores: the N read-modify-writes are independent (bricks NC apart) and overlap.
template <int N, int REC>
__device__ __forceinline__ void march_chains(uint32_t (&adr)[2], int& k, int kend, uint32_t ring_lane, int lane_wx, int lane_wy)
{
    if (k >= kend) return;
    // software pipeline: the record of iteration k + 1 is loaded while the ring words of iteration k are read, updated
    // and written, so only the ring's load -> fma -> store chain separates two markers of a brick.  On return adr[] points at
    // the first record that was not consumed.
    int2 zo[N];
    double wx[N], wy[N];
    double2 wz[N];
#pragma unroll
    for (int c = 0; c < N; ++c)
    {
        zo[c] = lds_v2s32(adr[c]);
        wz[c] = lds_v2f64(adr[c] + 16);
        wx[c] = lds_f64(adr[c] + lane_wx);
        wy[c] = lds_f64(adr[c] + lane_wy);
        adr[c] += REC;
    }
    for (; k < kend; ++k)
    {
        double a0[N], a1[N];
        uint32_t p0[N], p1[N];
        double wa[N], wb[N];
#pragma unroll
        for (int c = 0; c < N; ++c)
        {
            p0[c] = ring_lane + zo[c].x;
            p1[c] = ring_lane + zo[c].y;
            a0[c] = lds_f64(p0[c]);
            a1[c] = lds_f64(p1[c]);
            const double wxy = wx[c] * wy[c];
            wa[c] = wxy * wz[c].x;
            wb[c] = wxy * wz[c].y;
        }
        // (the loads after a chain's last record are wasted wavefronts, yet guarding them -- by a branch or by predicated
        // loads -- measured 6 % slower)
#pragma unroll
        for (int c = 0; c < N; ++c)
        {
            zo[c] = lds_v2s32(adr[c]);
            wz[c] = lds_v2f64(adr[c] + 16);
            wx[c] = lds_f64(adr[c] + lane_wx);
            wy[c] = lds_f64(adr[c] + lane_wy);
            adr[c] += REC;
        }
#pragma unroll
        for (int c = 0; c < N; ++c)
        {
            sts_f64(p0[c], a0[c] + wa[c]);
            sts_f64(p1[c], a1[c] + wb[c]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < N; ++c) adr[c] -= REC; // the record after the last one was not consumed
}

template <int K>
__global__ void __launch_bounds__(MarchCfg<K>::NT, 1)
    spread_march_kernel(const __grid_constant__ TileParams tp, const __grid_constant__ TmaMapSet maps, const SpreadArgs args)
{
    using C = MarchCfg<K>;
    constexpr int W = C::W, M = C::M, R = C::R, XO = C::XO, RX = C::RX, PLANE = C::PLANE, FZ = C::FZ, NRING = C::NRING;
    constexpr int NC = C::NC, NCW = C::NCW, REC = C::REC, ZO_BYTES = C::ZO_BYTES;
    constexpr int NPROD = 32 * C::NPW;
    constexpr int PLANE_B = PLANE * 8;
    static_assert(PLANE <= 2048, "g_march_zeros holds one plane");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const ringb = smem_raw;                                  // [NRING][R][RX] doubles
    // FAST4: a record that must not be spread here (a chain that has run out, a stencil left to the fix-up) points at a
    // SINK behind the ring instead of being branched around; record `cap` of each buffer is such a dummy
    constexpr int SINK_OFF = NRING * PLANE_B; // (relative to the ring)
    constexpr int SINK_B = C::SINK_B;
    unsigned char* const recb = smem_raw + (size_t)NRING * PLANE_B + SINK_B; // [2][cap + 1][REC]
    const int cap1 = args.cap + 1;
    __shared__ int s_pre[2][MARCH_BR * MARCH_BR + 1]; // markers before brick p of the layer (row-major: p = row * 8 + brick in row)
    __shared__ int s_first[2][MARCH_BR * MARCH_BR];   // sorted position of the brick's first marker
    __shared__ int s_desc[2][8];                      // step: layer, window offset, markers in the window, pre/first buffer, first window?, end?
    __shared__ int s_tot[MARCH_MAX_LAYERS + 1];       // markers per layer (decides which planes are flushed)
    __shared__ int s_any;
    __shared__ __align__(8) uint64_t s_zbar; // completion of the zero refills of the flusher

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NCONS = C::NCONS, CPW = C::CPW;
    const int role = warp < NCONS ? 0 : (warp == NCONS ? 1 : 2); // consumer / flusher / producer
    const int ptid = tid - 32 * (NCONS + 1);                     // producer thread index

    // ---- the work item in hand (set by the item loop below; the lambdas read it)
    int col[3] = { 0, 0, 0 }; // march tile
    int a = 0;                // component
    int tz0 = 0, NL = 0;      // first marker tile in z, brick layers of the chunk
    int bx0 = 0, by0 = 0, zq0 = 0; // smem column 0 / row 0 / plane 0 of the accumulator in pp coordinates
    bool use_tma = false;
    __shared__ int s_item;

    // tile selection (halo overlap): is tile (txi, tyi, tz) of this column part of this launch?
    auto tile_on = [&](int txi, int tyi, int tz) -> bool {
        const int tx = 2 * col[0] + txi, ty = 2 * col[1] + tyi;
        if (tx >= tp.nt[0] || ty >= tp.nt[1]) return false;
        if (!args.part) return true;
        const bool in = tx >= args.sel_lo[0] && tx <= args.sel_hi[0] && ty >= args.sel_lo[1] && ty <= args.sel_hi[1] &&
                        tz >= args.sel_lo[2] && tz <= args.sel_hi[2];
        return (args.part == 1) == in;
    };
    auto tile_b0 = [&](int txi, int tyi, int tz) -> int {
        const int tx = 2 * col[0] + txi, ty = 2 * col[1] + tyi;
        return tp.brick_base + ((tz * tp.nt[1] + ty) * tp.nt[0] + tx) * 64;
    };
    if (tid == 0)
    {
        mbar_init(&s_zbar, 1);
        mbar_fence_init();
    }
    // ---- zero the ring
    {
        double2* r2 = reinterpret_cast<double2*>(ringb);
        for (int q = tid; q < (NRING * PLANE_B + SINK_B) / 16; q += C::NT) r2[q] = make_double2(0.0, 0.0);
        if (C::FAST4 && tid < 2 * (REC / 4)) // the dummy records: offsets = the sink, weights = 0
        {
            int* dr = reinterpret_cast<int*>(recb + ((size_t)(tid / (REC / 4)) * cap1 + args.cap) * REC);
            const int wd = tid % (REC / 4);
            dr[wd] = (wd == 0 || wd == 1 || wd == 8 || wd == 9) ? SINK_OFF : 0;
        }
        fence_proxy_async_smem();
    }

    // =========================================================================================
    // producer: counts of a layer, then one thread per marker
    // =========================================================================================
    int p_layer = -1, p_off = 0, p_total = 0, p_lb = 1; // (uniform over the producer threads)
    int n_s = 0, n_m = 0, n_e = 0;                      // prefetched segment offsets of the lane's two bricks (next layer)
    bool n_ok = false, prefetch_pending = false;
    // lane l of producer warp 0 holds bricks p = 2 l, 2 l + 1 (same row, same tile, consecutive ids)
    auto counts_fetch = [&](int layer) {
        n_ok = false;
        if (ptid < 32 && layer < NL)
        {
            const int row = ptid >> 2, i = (2 * ptid) & 7;
            const int tz = tz0 + layer / TILE_BRICKS, lz = layer % TILE_BRICKS;
            const int txi = i >> 2, tyi = row >> 2;
            if (tile_on(txi, tyi, tz))
            {
                const int b = tile_b0(txi, tyi, tz) + 16 * lz + 4 * (row & 3) + (i & 3);
                n_s = __ldg(&args.brick_start[b]);
                n_m = __ldg(&args.brick_start[b + 1]);
                n_e = __ldg(&args.brick_start[b + 2]);
                n_ok = true;
            }
        }
    };
    auto counts_publish = [&](int lb) -> void {
        // (producer warp 0) exclusive scan of the 64 counts, two per lane
        if (ptid < 32)
        {
            int c0 = n_ok ? n_m - n_s : 0, c1 = n_ok ? n_e - n_m : 0;
            if (args.dense_thresh > 0)
            {
                if (c0 > args.dense_thresh) c0 = 0; // a dense brick: spread_dense_kernel's
                if (c1 > args.dense_thresh) c1 = 0;
            }
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - c0 - c1;
            s_pre[lb][2 * ptid + 1] = excl + c0;
            s_pre[lb][2 * ptid + 2] = incl;
            if (ptid == 0) s_pre[lb][0] = 0;
            s_first[lb][2 * ptid] = n_s;
            s_first[lb][2 * ptid + 1] = n_m;
        }
    };
    // prepares step t + 1 into buffers (nb = step parity): descriptor, (new layer: counts), records
    auto produce = [&](int nb) {
        int first_win = 0;
        if (p_layer >= 0 && p_off + args.cap < p_total)
            p_off += args.cap;
        else
        {
            ++p_layer;
            p_off = 0;
            p_lb ^= 1;
            first_win = 1;
            if (p_layer < NL)
            {
                counts_publish(p_lb);
                counts_fetch(p_layer + 1);
                prefetch_pending = true;
                named_bar_sync(2, NPROD);
                p_total = s_pre[p_lb][MARCH_BR * MARCH_BR];
            }
            else
                p_total = 0;
        }
        const int cnt = min(args.cap, p_total - p_off);
        if (ptid == 0)
        {
            s_desc[nb][0] = p_layer;
            s_desc[nb][1] = p_off;
            s_desc[nb][2] = cnt;
            s_desc[nb][3] = p_lb;
            s_desc[nb][4] = first_win;
            s_desc[nb][5] = p_layer >= NL ? 1 : 0;
            if (first_win && p_layer <= MARCH_MAX_LAYERS) s_tot[p_layer] = p_total;
        }
        if (p_layer >= NL) return;
        unsigned char* const rb = recb + (size_t)nb * cap1 * REC;
        const int* pre = s_pre[p_lb];
        const int zlo = BRICK * p_layer; // the layer's first plane (relative to plane 0 of the chunk)
        for (int slot = ptid; slot < cnt; slot += NPROD)
        {
            const int m = p_off + slot;
            int p = 0; // last p with pre[p] <= m
#pragma unroll
            for (int step = 32; step >= 1; step >>= 1)
                if (pre[p + step] <= m) p += step;
            const int i = s_first[p_lb][p] + (m - pre[p]);
            double xs[3], xr[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                xs[d] = __ldg(&args.X[d * args.x_stride + i]);
                xr[d] = args.Xraw ? __ldg(&args.Xraw[d * args.x_stride + i]) : xs[d];
            }
            const long long row = args.src ? (long long)__ldg(&args.src[i]) : (long long)i;
            const double fv = __ldg(&args.V[tp.comp[a].vcol * args.v_cstride + row * args.v_istride]) * tp.inv_vol;
            unsigned char* const r = rb + (size_t)slot * REC;
            double* const wr = reinterpret_cast<double*>(r + (C::FAST4 ? 64 : ZO_BYTES));
            int r0[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                double w[W];
                int l;
                stencil_1d<K>(xs[d], xr[d], tp.xl[d][tp.comp[a].var[d]], tp.dx[d], l, w, d == tp.comp[a].axis);
                r0[d] = l + tp.G - (d == 0 ? bx0 : d == 1 ? by0 : zq0);
#pragma unroll
                for (int j = 0; j < W; ++j)
                {
                    if (d < 2)
                        wr[d * W + j] = w[j];
                    else if (C::FAST4)
                        *reinterpret_cast<double*>(r + 32 * (j & 1) + 16 + 8 * (j >> 1)) = w[j] * fv;
                    else
                        wr[2 * W + j] = w[j] * fv;
                }
            }
            // the stencil must stay inside the footprint of the marker's BRICK (what the chains and row colours rely on) and
            // inside the block
            const int fx = MARCH_CELLS * col[0] + BRICK * (p & 7) - M - bx0, fy = MARCH_CELLS * col[1] + BRICK * (p >> 3) - M - by0;
            const bool fits = r0[0] >= max(fx, 0) && r0[0] + W <= min(fx + FZ, RX) && r0[1] >= max(fy, 0) && r0[1] + W <= min(fy + FZ, R) &&
                              r0[2] >= zlo && r0[2] + W <= zlo + FZ;
            const int xy = 8 * (r0[1] * RX + r0[0]);
            int zoff[W];
#pragma unroll
            for (int j = 0; j < W; ++j)
            {
                // a stencil that does not fit: FAST4 adds it to the sink (branch-free consumer), else it is skipped
                const int o = fits ? ((r0[2] + j) % NRING) * PLANE_B + xy : (C::FAST4 ? SINK_OFF : -1);
                if (C::FAST4)
                    zoff[j] = o;
                else
                    reinterpret_cast<int*>(r)[j] = o;
            }
            if (C::FAST4) // planes (g, g + 2) of a lane group sit side by side: one 8-byte store each (4-byte stores of 32 records
            {             // 144 bytes apart hit 8 banks four times over)
                *reinterpret_cast<int2*>(r) = make_int2(zoff[0], zoff[2]);
                *reinterpret_cast<int2*>(r + 32) = make_int2(zoff[1], zoff[3]);
            }
            if (!fits) flag_exception(args, i, a);
        }
        // the marker data of the NEXT layer (its segment offsets have arrived by now) on its way into L2
        if (prefetch_pending && ptid < 32 && n_ok && n_e > n_s)
        {
            const long long vrow = tp.comp[a].vcol * args.v_cstride;
            for (int i = n_s & ~15; i < n_e; i += 16) // 128-byte lines
            {
#pragma unroll
                for (int d = 0; d < 3; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(&args.X[d * args.x_stride + i]));
                if (!args.src) asm volatile("prefetch.global.L2 [%0];" ::"l"(&args.V[vrow + (long long)i * args.v_istride]));
            }
        }
        prefetch_pending = false;
    };

    // =========================================================================================
    // consumer: rows of bricks, NC colours of rows, chains of bricks NC apart inside a row
    // =========================================================================================
    auto consume = [&](int b, int lb, int off, int cnt) {
        const unsigned char* const rb = recb + (size_t)b * cap1 * REC;
        const int* pre = s_pre[lb];
        // this lane's stencil point(s)
        int lxy[C::NSLOT], lz4[C::NSLOT], lwx[C::NSLOT], lwy[C::NSLOT], lwz[C::NSLOT];
#pragma unroll
        for (int s = 0; s < C::NSLOT; ++s)
        {
            const int q = min(lane + 32 * s, C::NPTS - 1);
            const int ix = q % W, iy = (q / W) % W, iz = q / (W * W);
            lxy[s] = 8 * (iy * RX + ix);
            lz4[s] = 4 * iz;
            lwx[s] = ZO_BYTES + 8 * ix;
            lwy[s] = ZO_BYTES + 8 * (W + iy);
            lwz[s] = ZO_BYTES + 8 * (2 * W + iz);
        }
        const int g4 = lane >> 4; // FAST4: z pair (g4, g4 + 2)
        const uint32_t rec_s = smem_u32(rb);
        const uint32_t ring_lane = smem_u32(ringb) + (uint32_t)lxy[0];
        // FAST4: this lane's wx / wy relative to its 32-byte block of the record
        const int lane_wx = 64 + 8 * (lane & 3) - 32 * g4, lane_wy = 96 + 8 * ((lane >> 2) & 3) - 32 * g4;
        const int rslot = warp / C::WPR, half = warp % C::WPR; // which row of the colour, which half of its chains
        for (int ph = 0; ph < NC; ++ph)
        {
            const int row = ph + NC * rslot;
            if (row < MARCH_BR)
            {
#pragma unroll 1
                for (int sp = 0; sp < NC; ++sp)
                {
                    int cur[CPW], end[CPW];
                    int longest = 0;
#pragma unroll
                    for (int c = 0; c < CPW; ++c)
                    {
                        const int i = sp + NC * (CPW * half + c);
                        cur[c] = end[c] = 0;
                        if (i < MARCH_BR)
                        {
                            const int p = row * MARCH_BR + i;
                            cur[c] = max(pre[p], off) - off;
                            end[c] = min(pre[p + 1], off + cnt) - off;
                            longest = max(longest, end[c] - cur[c]);
                        }
                    }
                    if constexpr (C::FAST4)
                    {
                        // two chains side by side while both last, then the longer one alone: nothing idles and nothing
                        // branches inside a body.  (The order of the additions at a grid point is the order inside its
                        // brick's chain: same-sub-phase bricks are disjoint.)
                        static_assert(CPW == 2, "two chains per consumer warp");
                        uint32_t adr[2];
                        int len[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                        {
                            adr[c] = rec_s + (uint32_t)cur[c] * REC + 32 * g4;
                            len[c] = max(end[c] - cur[c], 0);
                        }
                        if (len[0] < len[1])
                        {
                            const int tl = len[0];
                            len[0] = len[1];
                            len[1] = tl;
                            const uint32_t ta = adr[0];
                            adr[0] = adr[1];
                            adr[1] = ta;
                        }
                        int k = 0;
                        march_chains<2, C::REC>(adr, k, len[1], ring_lane, lane_wx, lane_wy);
                        march_chains<1, C::REC>(adr, k, len[0], ring_lane, lane_wx, lane_wy);
                        // the next sub-phase of this row touches the neighbouring bricks: wait for the other half of the row
                        // (of the other half's bricks only brick 4, the upper half's first, is a neighbour of a brick of the next
                        // sub-phase, and that is the lower half's brick 3: the upper half announces, the lower half waits)
                        if (sp + 1 < NC)
                        {
                            if (IBK_MARCH_SUBARRIVE && half == 1)
                                named_bar_arrive(4 + rslot, 32 * C::WPR);
                            else
                                named_bar_sync(4 + rslot, 32 * C::WPR);
                        }
                    }
                    if constexpr (!C::FAST4)
                    {
                        for (int k = 0; k < longest; ++k)
                        {
#pragma unroll
                            for (int c = 0; c < CPW; ++c)
                            {
                                if (cur[c] + k >= end[c]) continue;
                                const unsigned char* r = rb + (size_t)(cur[c] + k) * REC;
                                if (*reinterpret_cast<const int*>(r) < 0) continue; // does not fit: the fix-up's
                                double wv[C::NSLOT], av[C::NSLOT];
                                int ad[C::NSLOT];
#pragma unroll
                                for (int s = 0; s < C::NSLOT; ++s)
                                {
                                    ad[s] = *reinterpret_cast<const int*>(r + lz4[s]) + lxy[s];
                                    wv[s] = (*reinterpret_cast<const double*>(r + lwx[s]) * *reinterpret_cast<const double*>(r + lwy[s])) *
                                            *reinterpret_cast<const double*>(r + lwz[s]);
                                }
#pragma unroll
                                for (int s = 0; s < C::NSLOT; ++s)
                                    if (C::NPTS % 32 == 0 || lane + 32 * s < C::NPTS) av[s] = *reinterpret_cast<const double*>(ringb + ad[s]);
#pragma unroll
                                for (int s = 0; s < C::NSLOT; ++s)
                                    if (C::NPTS % 32 == 0 || lane + 32 * s < C::NPTS) *reinterpret_cast<double*>(ringb + ad[s]) = av[s] + wv[s];
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            else if constexpr (C::FAST4)
            {
                // (a row slot without a row in this colour: FAST4 has 8 rows in 2 colours of 4, so this does not happen)
            }
            if (ph + 1 < NC)
            {
                if constexpr (C::FAST4 && IBK_MARCH_NBAR)
                {
                    // row 2 s + 1 overlaps rows 2 s (this pair's own) and 2 s + 2 (the next pair's) only: the pair tells the pair
                    // before it that its even row is done and waits for the pair after it, not for all four rows
                    if (rslot > 0) named_bar_arrive(8 + rslot, 64 * C::WPR);
                    if (rslot + 1 < NCW)
                        named_bar_sync(8 + rslot + 1, 64 * C::WPR);
                    else
                        named_bar_sync(8, 32 * C::WPR); // (the last pair: its own two warps)
                }
                else
                    named_bar_sync(1, 32 * NCONS);
            }
        }
        fence_proxy_async_smem(); // this thread's writes to the ring -> visible to the TMA stores of the flusher
    };

    // =========================================================================================
    // flusher: planes [qlo, qhi) are final: f += plane (TMA reducing store), then zero them for reuse
    // =========================================================================================
    uint32_t zphase = 0;
    auto flush = [&](int qlo, int qhi) {
        const int x0 = bx0 - tp.comp[a].pp0[0], y0 = by0 - tp.comp[a].pp0[1];
        // which of the planes were touched: plane q by the layers s with 4 s <= q < 4 s + FZ that held markers
        unsigned dirty = 0;
        for (int s = max(0, (qlo - FZ + BRICK) / BRICK); s <= min((qhi - 1) / BRICK, NL - 1); ++s)
            if (s_tot[s] > 0)
                for (int q = max(qlo, BRICK * s); q < min(qhi, BRICK * s + FZ); ++q) dirty |= 1u << (q - qlo);
        if (!dirty) return;
        for (int q = qlo; q < qhi; ++q)
        {
            if (!((dirty >> (q - qlo)) & 1u)) continue;
            const int gz = zq0 + q - tp.comp[a].pp0[2];
            if (gz < 0 || gz >= tp.comp[a].n[2]) continue;
            const double* pl = reinterpret_cast<const double*>(ringb + (size_t)(q % NRING) * PLANE_B);
            if (use_tma)
            {
                if (lane == 0) tma_reduce_add_3d(&maps.m[a], pl, x0, y0, gz);
            }
            else
            {
                // the array is not addressable by TMA here (alignment, or the block starts before the array): plain
                // read-modify-write by the lanes, clipped to the array.  One CTA per grid point and launch: fixed order.
                for (int y = 0; y < R; ++y)
                {
                    const int gy = y0 + y;
                    if (gy < 0 || gy >= tp.comp[a].n[1]) continue;
                    double* grow = tp.comp[a].ptr + ((long long)gz * tp.comp[a].n[1] + gy) * tp.comp[a].pitch;
                    for (int x = lane; x < RX; x += 32)
                    {
                        const int gx = x0 + x;
                        const double v = pl[y * RX + x];
                        if (gx >= 0 && gx < tp.comp[a].n[0] && v != 0.0) grow[gx] += v;
                    }
                }
            }
        }
        if (use_tma)
        {
            // zeros for the next use of the slots, by bulk copies from a plane of zeros (L2-resident) once the stores have
            // read the planes; the TMA unit does the work, this warp only waits
            if (lane == 0)
            {
                tma_store_commit_and_wait_read();
                mbar_expect_tx(&s_zbar, (uint32_t)(__popc(dirty) * PLANE_B));
                for (int q = qlo; q < qhi; ++q)
                    if ((dirty >> (q - qlo)) & 1u) bulk_copy_g2s(ringb + (size_t)(q % NRING) * PLANE_B, g_march_zeros, PLANE_B, &s_zbar);
            }
            mbar_wait(&s_zbar, zphase);
            zphase ^= 1u;
            return;
        }
        __syncwarp();
        for (int q = qlo; q < qhi; ++q)
        {
            if (!((dirty >> (q - qlo)) & 1u)) continue;
            double2* pl = reinterpret_cast<double2*>(ringb + (size_t)(q % NRING) * PLANE_B);
            for (int e = lane; e < PLANE / 2; e += 32) pl[e] = make_double2(0.0, 0.0);
        }
        fence_proxy_async_smem();
    };

    // ---- one-time set-up is done; the item loop
    __syncthreads();
    const int ncomp = tp.ncomp;
    for (;;)
    {
        if (tid == 0) s_item = atomicAdd(args.work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= args.item_base[8]) break;
        // decode: colour, then (position, component) with the component fastest (the three components of a tile run side by side
        // on three SMs and share the marker data in L2)
        int c = 0;
        while (item >= args.item_base[c + 1]) ++c;
        int ntc[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) ntc[d] = (args.nm[d] - ((c >> d) & 1) + 1) / 2;
        {
            int r = item - args.item_base[c];
            a = r % ncomp;
            r /= ncomp;
            col[0] = 2 * (r % ntc[0]) + (c & 1);
            r /= ntc[0];
            col[1] = 2 * (r % ntc[1]) + ((c >> 1) & 1);
            col[2] = 2 * (r / ntc[1]) + ((c >> 2) & 1);
        }
        tz0 = col[2] * args.chunk_tiles;
        const int ntz = min(args.chunk_tiles, tp.nt[2] - tz0);
        NL = ntz * TILE_BRICKS;
        int* const my_flag = args.done_flags + (((long long)col[2] * args.nm[1] + col[1]) * args.nm[0] + col[0]) * ncomp + a;
        // anything to spread in this chunk?
        if (tid == 0) s_any = 0;
        __syncthreads();
        if (tid < 4 * ntz)
        {
            const int tz = tz0 + tid / 4, txi = tid & 1, tyi = (tid >> 1) & 1;
            if (tile_on(txi, tyi, tz))
            {
                const int b0 = tile_b0(txi, tyi, tz);
                if (__ldg(&args.brick_start[b0 + 64]) > __ldg(&args.brick_start[b0])) s_any = 1;
            }
        }
        __syncthreads();
        if (!s_any)
        {
            if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], 1;" ::"l"(my_flag) : "memory");
            continue;
        }
        // geometry of the accumulator
        bx0 = MARCH_CELLS * col[0] - M - XO;
        by0 = MARCH_CELLS * col[1] - M;
        zq0 = TILE * tz0 - M;
        use_tma = (args.tma_mask >> a) & 1u;
        if (bx0 < tp.comp[a].pp0[0])
        {
            if (args.clip_free) bx0 = tp.comp[a].pp0[0]; // nothing reaches the points before the array: start the block at its first element
            else use_tma = false;                        // a TMA store with a negative start coordinate traps (measured)
        }
        if (by0 < tp.comp[a].pp0[1])
        {
            if (args.clip_free) by0 = tp.comp[a].pp0[1];
            else use_tma = false;
        }
        // The neighbouring march tiles of a LOWER colour overlap this one's halo and must have been added to f completely
        // before this item adds anything (fixed order of the additions at every grid point).  They hold smaller tickets, so
        // they are running or done: no deadlock.
        if (tid < 27 && tid != 13)
        {
            const int n0 = col[0] + tid % 3 - 1, n1 = col[1] + (tid / 3) % 3 - 1, n2 = col[2] + tid / 9 - 1;
            if (n0 >= 0 && n0 < args.nm[0] && n1 >= 0 && n1 < args.nm[1] && n2 >= 0 && n2 < args.nm[2] &&
                (n0 & 1) + 2 * (n1 & 1) + 4 * (n2 & 1) < c)
            {
                const int* fl = args.done_flags + (((long long)n2 * args.nm[1] + n1) * args.nm[0] + n0) * ncomp + a;
                int v;
                do
                {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
                    if (!v) __nanosleep(200);
                } while (!v);
            }
        }
        // ---- the march
        p_layer = -1;
        p_off = 0;
        p_total = 0;
        p_lb = 1;
        prefetch_pending = false;
        if (role == 2)
        {
            counts_fetch(0);
            produce(0);
        }
        __syncthreads();
        for (int t = 0;; ++t)
        {
            const int b = t & 1;
            const int layer = s_desc[b][0], off = s_desc[b][1], cnt = s_desc[b][2], lb = s_desc[b][3], first_win = s_desc[b][4],
                      is_end = s_desc[b][5];
            if (role == 2)
            {
                if (!is_end) produce(b ^ 1);
            }
            else if (role == 0)
            {
                if (!is_end && cnt > 0) consume(b, lb, off, cnt);
            }
            else
            {
                if (is_end)
                    flush(BRICK * (NL - 1), BRICK * NL + 2 * M);
                else if (first_win && layer > 0)
                    flush(BRICK * (layer - 1), BRICK * layer);
            }
            if (is_end) break;
            __syncthreads();
        }
        // the item is complete when its reducing stores are: then the flag (release) lets the neighbours of higher colours go
        if (role == 1)
        {
            __threadfence(); // (the plain write-out of a block TMA cannot address)
            __syncwarp();
            if (lane == 0)
            {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], 1;" ::"l"(my_flag) : "memory");
            }
        }
    }
}

// =============================================================================================
// tile kernel (2D, and the 8-point kernel in 3D)
// =============================================================================================
// Brick colouring of a tile, worked out at compile time: the bricks of a tile in colour-major order
// (colour = brick index mod NC per dimension; NC bricks apart, two footprints of BRICK + 2M cells are disjoint).
template <int NDIM, int NC>
struct BrickColouring
{
    static constexpr int NBT = TILE_BRICKS;
    static constexpr int NB = (NDIM == 3) ? NBT * NBT * NBT : NBT * NBT;
    static constexpr int NCOL = (NDIM == 3) ? NC * NC * NC : NC * NC;
    int start[NCOL + 1];      // first colour-order position of each colour
    unsigned char order[NB];  // colour-order position -> brick-in-tile id (the binning's order, x fastest)
    unsigned char colour[NB]; // colour-order position -> colour
    constexpr BrickColouring() : start{}, order{}, colour{}
    {
        int pos = 0;
        for (int c = 0; c < NCOL; ++c)
        {
            start[c] = pos;
            const int c0 = c % NC, c1 = (c / NC) % NC, c2 = c / (NC * NC);
            for (int q = 0; q < NB; ++q)
            {
                const int lx = q % NBT, ly = (q / NBT) % NBT, lz = q / (NBT * NBT);
                if (lx % NC == c0 && ly % NC == c1 && (NDIM == 2 || lz % NC == c2))
                {
                    order[pos] = (unsigned char)q;
                    colour[pos] = (unsigned char)c;
                    ++pos;
                }
            }
        }
        start[NCOL] = pos;
    }
};
template <int NDIM, int NC>
__constant__ BrickColouring<NDIM, NC> c_colouring = BrickColouring<NDIM, NC>(); // uniform reads (start[], one colour[])
template <int NDIM, int NC>
__device__ const BrickColouring<NDIM, NC> d_colouring = BrickColouring<NDIM, NC>(); // per-lane reads (order[])

//  * One CTA takes one marker tile (16^ndim cells = 4^ndim bricks, one contiguous run of the sorted markers) and ONE
//    component, and accumulates the full stencils of its markers into a shared-memory block of (16 + 2M)^ndim points.
//  * The blocks of two tiles whose indices differ by 2 in some dimension are disjoint: 2^ndim launches (colours).
//  * Inside the CTA one warp takes one brick at a time and walks its markers in storage order with the 32 lanes spread
//    over the stencil points; the bricks are visited colour by colour with a CTA barrier between colours.
//  * 1-D weights are evaluated one thread per (marker, dimension) for a window of markers and parked in shared memory.
template <int NDIM, int K>
__global__ void __launch_bounds__(SPREAD_THREADS, (KTraits<K>::M <= 2) ? 3 : (KTraits<K>::M <= 3 ? 2 : 1))
    spread_tile_kernel(const __grid_constant__ TileParams tp, const __grid_constant__ TmaMapSet maps, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int R = TILE + 2 * M; // haloed block edge
    constexpr int NT = SPREAD_THREADS;
    constexpr int NWARPS = NT / 32;
    // TMA boxes of 8-byte elements must start on an even x coordinate and have an even x extent (16 bytes):
    // the block gets XO spare columns on the left and is RX wide in x.
    constexpr int XO = M & 1;
    constexpr int RX = (R + XO + 1) & ~1;
    constexpr int RPTS = (NDIM == 3) ? R * R * RX : R * RX;
    constexpr int NC = (BRICK + 2 * M + BRICK - 1) / BRICK; // brick colours per dimension
    using Colouring = BrickColouring<NDIM, NC>;
    constexpr int NBRICKS = Colouring::NB;
    constexpr int NPTS = (NDIM == 3) ? W * W * W : W * W;
    constexpr int NSLOT = (NPTS + 31) / 32;
    constexpr int LD = NDIM - 1;
    static_assert(R <= 24 && R < 128, "origins are kept in bytes");
    const Colouring& bc = c_colouring<NDIM, NC>;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* acc = reinterpret_cast<double*>(smem_raw);            // [R][R][RX] (z, y, x)
    double* wgt = acc + RPTS;                                     // [cap][NDIM][W]  1-D weights (force folded in)
    // [2][cap] byte offset of the stencil origin inside the block (negative: the stencil does not fit), summed over the
    // dimensions by shared-memory atomics; double-buffered by window parity
    int* relb = reinterpret_cast<int*>(wgt + args.cap * NDIM * W);
    __shared__ int bfirst[NBRICKS];   // first marker of the brick at colour-order position p
    __shared__ int bpre[NBRICKS + 1]; // markers in the bricks before colour-order position p
    __shared__ int wsum[2];
    __shared__ int wcol[2][2];       // first / last colour present in the window (double-buffered by window parity)
    __shared__ int sbs[NBRICKS + 1]; // the tile's slice of brick_start
    __shared__ __align__(8) uint64_t tma_bar;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // which marker tile (of this launch's colour) and which component
    int t[3] = { 0, 0, 0 };
    {
        int r = (int)blockIdx.x;
        t[0] = 2 * (r % args.ntc[0]) + args.colour[0];
        r /= args.ntc[0];
        t[1] = 2 * (r % args.ntc[1]) + args.colour[1];
        if (NDIM == 3) t[2] = 2 * (r / args.ntc[1]) + args.colour[2];
    }
    if (args.part)
    {
        bool in = true;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) in = in && t[d] >= args.sel_lo[d] && t[d] <= args.sel_hi[d];
        if ((args.part == 1) != in) return;
    }
    const int a = blockIdx.y;
    const CompGeom& cg = tp.comp[a];
    const int tile = (NDIM == 3) ? (t[2] * tp.nt[1] + t[1]) * tp.nt[0] + t[0] : t[1] * tp.nt[0] + t[0];
    const int b0 = tp.brick_base + tile * NBRICKS;
    // one coalesced read of the tile's NBRICKS + 1 segment offsets, and (independent of it) the colour order
    int my_q = 0;
    if (threadIdx.x <= NBRICKS) sbs[threadIdx.x] = __ldg(&args.brick_start[b0 + threadIdx.x]);
    if (threadIdx.x < NBRICKS) my_q = __ldg(&d_colouring<NDIM, NC>.order[threadIdx.x]);
    for (int q = threadIdx.x; q < args.cap; q += NT) relb[q] = 0;
    __syncthreads();
    const int s0 = sbs[0], s1 = sbs[NBRICKS];
    if (s0 >= s1) return;

    int blo[3]; // pp coordinate of the block's first point (x: of the first of the XO spare columns)
    bool inside = true; // the block does not start before the array
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        blo[d] = TILE * t[d] - M - (d == 0 ? XO : 0);
        if (d < NDIM && blo[d] < cg.pp0[d])
        {
            if (args.clip_free) blo[d] = cg.pp0[d]; // nothing reaches the points before the array
            else inside = false;
        }
    }
    // With TMA the block starts as a copy of f (out-of-array points read as zero) and is stored back at the end
    // (out-of-array points dropped): `f += S[F]` without a separate zero / add pass and without atomics.
    // (Measured on B200: a TMA tensor STORE with a negative start coordinate traps, loads do not.  The store also writes
    // whole 16-byte units, i.e. one element of the row padding when n[0] is odd: harmless, nothing reads the padding.)
    const bool use_tma = ((args.tma_mask >> a) & 1u) && inside;
    if (use_tma && threadIdx.x == 0)
    {
        mbar_init(&tma_bar, 1);
        mbar_fence_init();
    }

    // marker ranges of the bricks in colour-major order, and their running count (two-warp scan)
    int my_cnt = 0, my_incl = 0;
    if (threadIdx.x < NBRICKS)
    {
        const int s = sbs[my_q], e = sbs[my_q + 1];
        bfirst[threadIdx.x] = s;
        my_cnt = e - s;
        if (args.dense_thresh > 0 && my_cnt > args.dense_thresh) my_cnt = 0; // a dense brick: spread_dense_kernel's
    }
    if (warp < 2)
    {
        my_incl = my_cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, my_incl, o);
            if (lane >= o) my_incl += v;
        }
        if (lane == 31) wsum[warp] = my_incl;
    }
    if (!use_tma)
        for (int q = threadIdx.x; q < RPTS; q += NT) acc[q] = 0.0;
    __syncthreads();
    if (use_tma && threadIdx.x == 0)
    {
        mbar_expect_tx(&tma_bar, (uint32_t)(RPTS * sizeof(double)));
        if (NDIM == 3)
            tma_load_3d(acc, &maps.m[a], &tma_bar, blo[0] - cg.pp0[0], blo[1] - cg.pp0[1], blo[2] - cg.pp0[2]);
        else
            tma_load_2d(acc, &maps.m[a], &tma_bar, blo[0] - cg.pp0[0], blo[1] - cg.pp0[1]);
    }
    if (threadIdx.x < NBRICKS)
    {
        const int before = (warp == 1) ? wsum[0] : 0;
        bpre[threadIdx.x + 1] = before + my_incl;
        if (threadIdx.x == 0) bpre[0] = 0;
    }

    // ---- The tile's markers are taken in windows of `cap` markers in COLOUR-MAJOR brick order.  Per window:
    // (A) all threads evaluate the 1-D stencils, one thread per (marker, dimension); (B) brick colour by brick
    // colour, each warp walks the markers of one brick with the 32 lanes spread over the stencil points.
    const double* Xp = args.X;
    const double* Xr = args.Xraw;
    const int G = tp.G;
    const double inv_vol = tp.inv_vol;
    const int vcol = cg.vcol;
    const int cap = args.cap;
    // phase-B role: this lane's stencil point(s), as offsets into the block and into the weight scratch
    int poffb[NSLOT], pw0[NSLOT], pw1[NSLOT], pw2[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
    {
        const int q = lane + 32 * s;
        const int ix = q % W, iy = (q / W) % W, iz = (NDIM == 3) ? q / (W * W) : 0;
        poffb[s] = 8 * ((iz * R + iy) * RX + ix);
        pw0[s] = ix;
        pw1[s] = W + iy;
        pw2[s] = 2 * W + iz;
    }
    __syncthreads();
    const int total = bpre[NBRICKS]; // without the dense bricks

    // stencil-evaluation task of this thread for a window: the global loads are issued by fetch() -- for window
    // w + 1 before the accumulation of window w starts, so their latency hides behind it -- and consumed by
    // evaluate() at the top of the window.
    int t_i, t_fp = 0; // (t_fp: first block coordinate of the marker's BRICK footprint along the task's dimension)
    double t_xs, t_xr, t_v;
    auto fetch = [&](int off, int par) {
        const int cnt = min(cap, total - off);
        const int tix = threadIdx.x;
        t_i = -1;
        if (tix >= cnt * NDIM) return;
        const int m = tix / NDIM, d = tix - m * NDIM;
        const int lp = off + m; // position in the tile's colour-major marker list
        int p = 0;              // its brick (colour-order position): last p with bpre[p] <= lp
#pragma unroll
        for (int step = NBRICKS / 2; step >= 1; step >>= 1)
            if (bpre[p + step] <= lp) p += step;
        if (d == 0 && m == 0) wcol[par][0] = bc.colour[p];
        if (d == 0 && m == cnt - 1) wcol[par][1] = bc.colour[p];
        const int i = bfirst[p] + (lp - bpre[p]);
        if (d == 0 && off > 0) relb[par * cap + m] = 0;
        t_i = i;
        {
            const int q = d_colouring<NDIM, NC>.order[p]; // brick in the tile, x fastest
            const int ld = (q >> (2 * d)) & 3;
            const int blo_d = (d == 0) ? blo[0] : (d == 1) ? blo[1] : blo[2];
            t_fp = TILE * t[d] + BRICK * ld - M - blo_d;
        }
        t_xs = __ldg(&Xp[d * args.x_stride + i]);
        t_xr = Xr ? __ldg(&Xr[d * args.x_stride + i]) : 0.0;
        t_v = 1.0;
        if (d == LD)
        {
            const long long row = args.src ? (long long)__ldg(&args.src[i]) : (long long)i;
            t_v = __ldg(&args.V[vcol * args.v_cstride + row * args.v_istride]);
        }
    };
    auto evaluate = [&](int par) {
        if (t_i < 0) return;
        const int tix = threadIdx.x;
        const int m = tix / NDIM, d = tix - m * NDIM;
        const double xl_s = tp.xl[d][cg.var[d]];
        const double dx_s = tp.dx[d];
        const int blo_s = (d == 0) ? blo[0] : (d == 1) ? blo[1] : blo[2];
        double w[W];
        int l;
        stencil_1d<K>(t_xs, Xr ? t_xr : t_xs, xl_s, dx_s, l, w, d == cg.axis);
        const int r0 = l + G - blo_s; // first stencil point relative to the block
        // inside the block AND inside the footprint of the marker's brick (what the brick colours rely on: a stencil that
        // left its brick's footprint -- positions moved since the binning -- goes to the fix-up instead of racing)
        const bool fits = r0 >= max(t_fp, 0) && r0 + W <= min(t_fp + BRICK + 2 * M, (d == 0) ? RX : R);
        const double scale = (d == LD) ? t_v * inv_vol : 1.0;
#pragma unroll
        for (int j = 0; j < W; ++j) wgt[(m * NDIM + d) * W + j] = w[j] * scale;
        const int stride_b = (d == 0) ? 8 : (d == 1) ? 8 * RX : 8 * RX * R; // bytes per point along d in the block
        atomicAdd(&relb[par * cap + m], fits ? r0 * stride_b : -(1 << 29));
    };

    fetch(0, 0);
    int par = 0;
    for (int off = 0; off < total; off += cap, par ^= 1)
    {
        const int cnt = min(cap, total - off);
        // ---- phase A
        evaluate(par);
        __syncthreads();
        if (use_tma && off == 0) mbar_wait(&tma_bar, 0); // the block holds f now
        const int col_lo = wcol[par][0], col_hi = wcol[par][1];
        if (off + cap < total) fetch(off + cap, par ^ 1);
        // ---- phase B, colour by colour (only the colours this window holds)
        for (int col = col_lo; col <= col_hi; ++col)
        {
            const int cend = bc.start[col + 1];
            for (int p = bc.start[col] + warp; p < cend; p += NWARPS)
            {
                const int p0 = bpre[p], p1 = bpre[p + 1];
                const int m0 = max(p0, off) - off, m1 = min(p1, off + cnt) - off;
                if (m0 >= m1) continue;
                // Per marker only the block's read-modify-write is serial: the origin of marker m + 1 is read ahead,
                // the weights are addressed by running pointers, all loads of a marker precede its stores.
                constexpr int NW = NDIM * W;
                const int* rp = relb + par * cap;
                const double* w0p = wgt + m0 * NW + pw0[0];
                const double* w1p = wgt + m0 * NW + pw1[0];
                char* const accb = reinterpret_cast<char*>(acc);
                int ab_next = rp[m0];
                for (int m = m0; m < m1; ++m, w0p += NW, w1p += NW)
                {
                    const int ab = ab_next;
                    if (m + 1 < m1) ab_next = rp[m + 1];
                    if (ab >= 0) // else: does not fit the block, left to the fix-up (warp-uniform)
                    {
                        double wv[NSLOT], av[NSLOT];
                        if constexpr (NDIM == 3 && (32 % (W * W)) == 0)
                        {
                            // the lane's (ix, iy) is the same in every slot: one xy product, one z weight per slot
                            const double wxy = w0p[0] * w1p[0];
#pragma unroll
                            for (int s = 0; s < NSLOT; ++s) wv[s] = wxy * w0p[pw2[s] - pw0[0]];
                        }
                        else
                        {
#pragma unroll
                            for (int s = 0; s < NSLOT; ++s)
                            {
                                if (NPTS % 32 != 0 && lane + 32 * s >= NPTS) continue;
                                wv[s] = w0p[pw0[s] - pw0[0]] * w0p[pw1[s] - pw0[0]];
                                if (NDIM == 3) wv[s] *= w0p[pw2[s] - pw0[0]];
                            }
                        }
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s)
                            if (NPTS % 32 == 0 || lane + 32 * s < NPTS) av[s] = *reinterpret_cast<double*>(accb + ab + poffb[s]);
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s)
                            if (NPTS % 32 == 0 || lane + 32 * s < NPTS) *reinterpret_cast<double*>(accb + ab + poffb[s]) = av[s] + wv[s];
                    }
                    else if (lane == 0)
                        flag_exception(args, bfirst[p] + (off + m - p0), a);
                    __syncwarp();
                }
            }
            __syncthreads();
        }
    }

    // ---- write-out.  TMA: the block (= old f + the spread values) is stored back, clipped to the array.
    if (use_tma)
    {
        fence_proxy_async_smem(); // this thread's generic-proxy writes to the block -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (NDIM == 3)
                tma_store_3d(&maps.m[a], acc, blo[0] - cg.pp0[0], blo[1] - cg.pp0[1], blo[2] - cg.pp0[2]);
            else
                tma_store_2d(&maps.m[a], acc, blo[0] - cg.pp0[0], blo[1] - cg.pp0[1]);
            tma_store_commit_and_wait_read(); // shared memory must stay alive until it has been read
        }
        return;
    }
    // Fallback (array not addressable by TMA, or the block starts before the array): f += block by plain loads and
    // stores, rows along x, dropping points outside the array.  Inside one launch a grid point belongs to at most one CTA
    // and the launches are ordered by the stream, so the order of the additions is fixed.
    {
        constexpr int ROWS = (NDIM == 3) ? R * R : R;
        const int gx0 = blo[0] - cg.pp0[0], gy0 = blo[1] - cg.pp0[1], gz0 = (NDIM == 3) ? blo[2] - cg.pp0[2] : 0;
        for (int row = warp; row < ROWS; row += NWARPS)
        {
            const int gj = gy0 + row % R, gk = gz0 + row / R;
            if (gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
            double* prow = cg.ptr + ((long long)gk * cg.n[1] + gj) * cg.pitch;
            if (lane < RX)
            {
                const int gi = gx0 + lane;
                const double v = acc[row * RX + lane];
                if (gi >= 0 && gi < cg.n[0] && v != 0.0) prow[gi] += v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense bricks (3D).  A structure puts tens of markers into a cell, hundreds into a brick; the
// tile kernels would walk them with ONE warp (a brick is the unit of their colouring).  Here a CTA takes one dense
// brick and one component: every warp holds the brick's footprint ((4 + 2M)^3 points; for M = 3 one z half of it) in
// registers, 12-20 points per lane, and adds every 8th (4th) batch of the brick's markers into it with zero-padded 1-D weight
// vectors -- branch-free, no shared-memory read-modify-write, no conflicts.  The eight partial footprints are then
// summed in a fixed tree and added to f.  Bricks NC apart have disjoint footprints: NC^3 launches (colours).
// Order of the additions: fixed by (batch order within a warp, tree over the warps, brick colour): reproducible.
// ---------------------------------------------------------------------------------------------
constexpr int DENSE_BATCH = 10; // markers per warp and stencil-evaluation round: 3 * 10 lanes busy

template <int K>
__global__ void __launch_bounds__(256, (KTraits<K>::M <= 2) ? 3 : 2)
    spread_dense_kernel(const __grid_constant__ TileParams tp, SpreadArgs args, const int* __restrict__ dense_list, int c0, int c1,
                        int c2)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int FP = BRICK + 2 * M;       // footprint edge: 6, 8 or 10 points
    constexpr int NC = (BRICK + 2 * M + BRICK - 1) / BRICK;
    constexpr int NXY = FP * FP;            // (x, y) columns of the footprint
    constexpr int PPL = (NXY + 31) / 32;    // columns per lane
    constexpr int ZSPLIT = (FP > 8) ? 2 : 1; // a 10^3 footprint is shared by a pair of warps (lower / upper z half)
    constexpr int ZN = FP / ZSPLIT;         // z planes per warp
    constexpr int NG = 8 / ZSPLIT;          // warp groups, each taking every NG-th batch of markers
    static_assert(FP % ZSPLIT == 0, "z planes split evenly");
    __shared__ double wpad[8][DENSE_BATCH][3][FP]; // zero-padded 1-D weights over the footprint, per warp
    __shared__ double red[NG / 2][FP * FP * FP];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp % NG, zh = warp / NG;
    const int b = __ldg(&dense_list[blockIdx.x]);
    const int relb = b - tp.brick_base;
    const int nbricks = tp.nt[0] * tp.nt[1] * tp.nt[2] * 64;
    if (relb < 0 || relb >= nbricks) return; // another patch's brick
    int gb[3];
    {
        int tile = relb >> 6;
        const int q = relb & 63;
        const int t0 = tile % tp.nt[0];
        tile /= tp.nt[0];
        gb[0] = 4 * t0 + (q & 3);
        gb[1] = 4 * (tile % tp.nt[1]) + ((q >> 2) & 3);
        gb[2] = 4 * (tile / tp.nt[1]) + (q >> 4);
    }
    if (gb[0] % NC != c0 || gb[1] % NC != c1 || gb[2] % NC != c2) return; // another colour's brick
    const int a = blockIdx.y;
    const CompGeom& cg = tp.comp[a];
    const int s = __ldg(&args.brick_start[b]), e = __ldg(&args.brick_start[b + 1]);
    int fo[3]; // pp coordinate of the footprint's first point
#pragma unroll
    for (int d = 0; d < 3; ++d) fo[d] = BRICK * gb[d] - M;

    // this lane's columns (x, y) of the footprint; a column index beyond NXY is parked on column 0 with weight 0
    int cx[PPL], cy[PPL];
    bool cok[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k)
    {
        const int c = lane + 32 * k;
        cok[k] = c < NXY;
        cx[k] = cok[k] ? c % FP : 0;
        cy[k] = cok[k] ? c / FP : 0;
    }
    double acc[PPL][ZN];
#pragma unroll
    for (int k = 0; k < PPL; ++k)
#pragma unroll
        for (int z = 0; z < ZN; ++z) acc[k][z] = 0.0;

    // the lane's stencil task (marker m of the batch, dimension d); its loads run one batch ahead
    const int tm = lane / 3, td = lane - 3 * tm;
    double nxs = 0.0, nxr = 0.0, nv = 1.0;
    auto fetch = [&](int first) {
        const int i = first + tm;
        if (lane < 3 * DENSE_BATCH && i < e)
        {
            nxs = __ldg(&args.X[td * args.x_stride + i]);
            nxr = args.Xraw ? __ldg(&args.Xraw[td * args.x_stride + i]) : nxs;
            if (td == 2)
            {
                const long long row = args.src ? (long long)__ldg(&args.src[i]) : (long long)i;
                nv = __ldg(&args.V[cg.vcol * args.v_cstride + row * args.v_istride]);
            }
        }
    };
    fetch(s + grp * DENSE_BATCH);
    for (int first = s + grp * DENSE_BATCH; first < e; first += NG * DENSE_BATCH)
    {
        const int nb = min(DENSE_BATCH, e - first);
        const double xs = nxs, xr = nxr, fv = nv;
        fetch(first + NG * DENSE_BATCH);
        if (lane < nb * 3)
        {
            const int m = tm, d = td;
            const int i = first + m;
            double w[W];
            int l;
            stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w, d == cg.axis);
            const int r0 = l + tp.G - fo[d];
            const bool fits = r0 >= 0 && r0 + W <= FP;
            const double scale = (d == 2) ? fv * tp.inv_vol : 1.0;
            double* wp = wpad[warp][m][d];
#pragma unroll
            for (int z = 0; z < FP; ++z) wp[z] = 0.0;
            if (fits)
            {
#pragma unroll
                for (int j = 0; j < W; ++j) wp[r0 + j] = w[j] * scale;
            }
            else if (zh == 0) // the whole (marker, component) goes to the fix-up; a zero factor removes it here
                flag_exception(args, i, a);
        }
        __syncwarp();
        for (int m = 0; m < nb; ++m)
        {
            const double* wp = &wpad[warp][m][0][0];
            double pxy[PPL];
#pragma unroll
            for (int k = 0; k < PPL; ++k) pxy[k] = cok[k] ? wp[cx[k]] * wp[FP + cy[k]] : 0.0;
#pragma unroll
            for (int z = 0; z < ZN; ++z)
            {
                const double wz = wp[2 * FP + zh * ZN + z];
#pragma unroll
                for (int k = 0; k < PPL; ++k) acc[k][z] += pxy[k] * wz;
            }
        }
        __syncwarp();
    }

    // fixed reduction tree over the warp groups (per z half): (g, g + NG/2), (g, g + NG/4), ...; point (x, y, z) at
    // (z * FP + y) * FP + x
    auto put = [&](double* dst) {
#pragma unroll
        for (int k = 0; k < PPL; ++k)
            if (cok[k])
#pragma unroll
                for (int z = 0; z < ZN; ++z) dst[((zh * ZN + z) * FP + cy[k]) * FP + cx[k]] = acc[k][z];
    };
    auto take = [&](const double* src) {
#pragma unroll
        for (int k = 0; k < PPL; ++k)
            if (cok[k])
#pragma unroll
                for (int z = 0; z < ZN; ++z) acc[k][z] += src[((zh * ZN + z) * FP + cy[k]) * FP + cx[k]];
    };
#pragma unroll
    for (int half = NG / 2; half >= 1; half >>= 1)
    {
        if (grp >= half && grp < 2 * half) put(red[grp - half]);
        __syncthreads();
        if (grp < half) take(red[grp]);
        __syncthreads();
    }
    if (grp == 0) put(red[0]);
    __syncthreads();
    // f += footprint, dropping the points outside the array (same-colour bricks are disjoint: plain read-modify-write)
    for (int pt = threadIdx.x; pt < FP * FP * FP; pt += 256)
    {
        const double v = red[0][pt];
        if (v == 0.0) continue;
        const int px = pt % FP, py = (pt / FP) % FP, pz = pt / (FP * FP);
        const int gi = fo[0] + px - cg.pp0[0], gj = fo[1] + py - cg.pp0[1], gk = fo[2] + pz - cg.pp0[2];
        if (gi < 0 || gi >= cg.n[0] || gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
        cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi] += v;
    }
}

// ---------------------------------------------------------------------------------------------
// Fix-up for the flagged (entry, component) pairs: ONE CTA walks the flag words in order; each flagged pair is spread by
// warp 0 with its whole stencil, clipped to the array only, lanes over the stencil points.  Order: sorted position, then
// component.  The words are cleared on the way, so the flags are all zero again afterwards.
// ---------------------------------------------------------------------------------------------
constexpr int FIXUP_THREADS = 256;
template <int NDIM, int K>
__global__ void __launch_bounds__(FIXUP_THREADS) spread_fixup_kernel(const __grid_constant__ TileParams tp, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int NPTS = (NDIM == 3) ? W * W * W : W * W;
    __shared__ int s_list[FIXUP_THREADS];
    __shared__ unsigned s_word[FIXUP_THREADS];
    __shared__ int s_wcnt[FIXUP_THREADS / 32 + 1];
    if (*args.exc_count <= 0) return; // (uniform)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwords = (args.n_entries + 3) / 4;
    for (int base = 0; base < nwords; base += FIXUP_THREADS)
    {
        const int wi = base + threadIdx.x;
        const unsigned w = wi < nwords ? args.exc_flags[wi] : 0u;
        // ordered compaction of the nonzero words of this chunk
        const unsigned bal = __ballot_sync(0xffffffffu, w != 0u);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < FIXUP_THREADS / 32; ++k)
        {
            if (k < warp) before += s_wcnt[k];
            total += s_wcnt[k];
        }
        if (w != 0u)
        {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            s_list[pos] = wi;
            s_word[pos] = w;
            args.exc_flags[wi] = 0u;
        }
        __syncthreads();
        if (warp == 0)
        {
            for (int e = 0; e < total; ++e)
            {
                const unsigned word = s_word[e];
                for (int bit = 0; bit < 32; ++bit)
                {
                    if (!((word >> bit) & 1u)) continue;
                    const int i = s_list[e] * 4 + (bit >> 3), a = bit & 7;
                    if (i >= args.n_entries || a >= tp.ncomp) continue;
                    const CompGeom& cg = tp.comp[a];
                    const long long row = args.src ? (long long)args.src[i] : (long long)i;
                    double w1[3][W];
                    int lo[3] = { 0, 0, 0 };
                    for (int d = 0; d < NDIM; ++d)
                    {
                        const double xs = args.X[d * args.x_stride + i];
                        const double xr = args.Xraw ? args.Xraw[d * args.x_stride + i] : xs;
                        int l;
                        stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w1[d], d == cg.axis);
                        lo[d] = l + tp.G;
                    }
                    const double f = args.V[cg.vcol * args.v_cstride + row * args.v_istride] * tp.inv_vol;
                    for (int q = lane; q < NPTS; q += 32)
                    {
                        const int ii = q % W, j = (q / W) % W, k = (NDIM == 3) ? q / (W * W) : 0;
                        const int gi = lo[0] + ii - cg.pp0[0], gj = lo[1] + j - cg.pp0[1], gk = (NDIM == 3) ? lo[2] + k - cg.pp0[2] : 0;
                        if (gi < 0 || gi >= cg.n[0] || gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
                        double wv = w1[0][ii] * w1[1][j];
                        if (NDIM == 3) wv *= w1[2][k];
                        cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi] += wv * f;
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *args.exc_count = 0;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int NDIM, int K>
static cudaError_t launch_spread_t(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    cudaError_t e;
    if (!bins.exc_flags || !bins.exc_count)
    {
        err = "spread: the bins carry no exception flags";
        return cudaErrorInvalidValue;
    }
    SpreadArgs args;
    std::memset(&args, 0, sizeof(args));
    args.brick_start = bins.brick_start;
    args.X = mv.X;
    args.Xraw = mv.Xraw;
    args.x_stride = mv.x_stride;
    args.V = mv.V;
    args.v_cstride = mv.v_cstride;
    args.v_istride = mv.v_istride;
    args.src = mv.src;
    args.exc_count = bins.exc_count;
    args.exc_flags = bins.exc_flags;
    args.n_entries = bins.n_entries;
    args.part = mv.part;
    for (int d = 0; d < 3; ++d)
    {
        args.sel_lo[d] = mv.sel_lo[d];
        args.sel_hi[d] = mv.sel_hi[d];
    }
    // does this patch hold any marker at all?
    bool any = bins.range_base.empty();
    for (size_t p = 0; p < bins.range_base.size(); ++p)
        if (bins.range_base[p] == tp.brick_base && bins.range_last[p] > bins.range_first[p]) any = true;
    if (!any) return cudaSuccess;
    static const bool no_tma = getenv("IBK_NO_TMA") != nullptr; // (debugging: every block takes the plain write-out)
    // 3D, reach of at most 2 cells: the march kernel.  Reach 3 (IB_6, IB_5, BSPLINE_5 / 6, ...): the march kernel works (the
    // whole parity suite passes with it) but its generic 6 x 6 x 6 consumer path is slower than the tile kernel
    // (measured, 2^23 uniform markers on 512^3, IB_6: march 27.3 ms, tile 14.3 ms), so those take the tile kernel, like the
    // 8-point kernel and 2D.
    constexpr bool MARCH = (NDIM == 3 && M <= 2);
    // the first block per dimension may start at the array's first element when no stencil reaches outside the arrays and
    // the shifted block stays clear of the next block of the same colour
    args.clip_free = mv.clip_free ? 1 : 0;
    for (int a = 0; a < tp.ncomp; ++a)
        for (int d = 0; d < NDIM; ++d)
            if (tp.comp[a].pp0[d] > (MARCH ? 2 * TILE : TILE) - 1 - 3 * M) args.clip_free = 0;
    TmaMapSet maps;
    std::memset(&maps, 0, sizeof(maps));
    args.tma_mask = 0;
    auto ffn = spread_fixup_kernel<NDIM, K>;

    // dense bricks first (their own kernel), then the tiles without them
    args.dense_thresh = 0;
    if constexpr (NDIM == 3 && M <= 3) // (a 12^3 footprint does not fit the dense kernel's static shared memory)
    {
        static const bool no_dense = getenv("IBK_NO_DENSE") != nullptr;
        if (bins.n_dense > 0 && !no_dense && mv.part != 1) // (with a tile selection the dense bricks go with the boundary part)
        {
            args.dense_thresh = DENSE_BRICK_MARKERS;
            constexpr int NCB = (BRICK + 2 * M + BRICK - 1) / BRICK;
            for (int c = 0; c < NCB * NCB * NCB; ++c)
            {
                spread_dense_kernel<K><<<dim3((unsigned)bins.n_dense, (unsigned)tp.ncomp), 256, 0, L.stream>>>(
                    tp, args, bins.dense_list, c % NCB, (c / NCB) % NCB, c / (NCB * NCB));
                L.launches++;
            }
        }
        else if (bins.n_dense > 0 && !no_dense)
            args.dense_thresh = DENSE_BRICK_MARKERS; // part 1: the dense bricks are (were) done with part 2 / 0
    }

    bool done = false;
    if constexpr (MARCH)
    {
        done = true;
        using C = MarchCfg<K>;
        // planes are added to f by TMA when it can address the array and the block starts on an even x coordinate
        for (int a = 0; a < tp.ncomp; ++a)
            if (!no_tma && ((tp.comp[a].pp0[0] + M + C::XO) % 2 == 0) && make_tensor_map(&maps.m[a], tp.comp[a], 3, C::RX, C::R, 1, 0))
                args.tma_mask |= (1u << a);
        constexpr size_t ring_bytes = (size_t)C::NRING * C::PLANE * sizeof(double);
        constexpr size_t budget = 232448 - 3072 - ring_bytes - C::SINK_B - 2 * C::REC; // (static shared memory: the per-layer tables)
        args.cap = (int)std::min<size_t>(384, budget / (2 * C::REC));
        const size_t smem = ring_bytes + C::SINK_B + 2 * (size_t)(args.cap + 1) * C::REC;
        // 8 marker tiles (128 cells, 32 brick layers) per chunk in z: measured best (4: 3.06 ms, 8: 2.89 ms on the C5 shard)
        args.chunk_tiles = MARCH_MAX_LAYERS / TILE_BRICKS;
        auto kfn = spread_march_kernel<K>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
        {
            err = "cudaFuncSetAttribute(spread march) failed";
            return e;
        }
        // persistent CTAs, one per SM (the ring fills the shared memory), taking (march tile, component) items by ticket in
        // colour-major order; an item waits for its neighbours of lower colours (done flags), so ONE launch does what took
        // one launch per colour, without their eight tails
        for (int d = 0; d < 3; ++d) args.nm[d] = d < 2 ? (tp.nt[d] + 1) / 2 : (tp.nt[2] + args.chunk_tiles - 1) / args.chunk_tiles;
        args.item_base[0] = 0;
        for (int c = 0; c < 8; ++c)
        {
            int ntiles = 1;
            for (int d = 0; d < 3; ++d) ntiles *= (args.nm[d] - ((c >> d) & 1) + 1) / 2;
            args.item_base[c + 1] = args.item_base[c] + ntiles * tp.ncomp;
        }
        const int n_items = args.item_base[8];
        const size_t n_flags = (size_t)args.nm[0] * args.nm[1] * args.nm[2] * tp.ncomp;
        if (n_items > 0)
        {
            if (!bins.march_sync || n_flags + 1 > bins.march_sync_capacity)
            {
                err = "spread march: the bins carry no work counter / done flags of that size";
                return cudaErrorInvalidValue;
            }
            args.work_counter = bins.march_sync;
            args.done_flags = bins.march_sync + 1;
            if ((e = cudaMemsetAsync(bins.march_sync, 0, sizeof(int) * (n_flags + 1), L.stream)) != cudaSuccess) return e;
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int grid = std::max(1, std::min(n_items, (sms > 0 ? sms : 148) - L.reserve_sms));
            kfn<<<grid, C::NT, smem, L.stream>>>(tp, maps, args);
            L.launches++;
        }
    }
    if (!done)
    {
        constexpr int R = TILE + 2 * M;
        constexpr int XO = M & 1;
        constexpr int RX = (R + XO + 1) & ~1;
        constexpr int RPTS = (NDIM == 3) ? R * R * RX : R * RX;
        // window size: as many markers as the shared memory left by the block allows at the target residency
        constexpr int target_ctas = (M <= 2) ? 3 : (M <= 3 ? 2 : 1);
        constexpr long long budget = 233472 / target_ctas - 1024 - 2304 - (long long)sizeof(double) * RPTS;
        constexpr int per_marker = (int)sizeof(double) * NDIM * W + 2 * (int)sizeof(int);
        constexpr int cap_fit = (int)(budget / per_marker);
        args.cap = std::max(32, std::min(256, cap_fit));
        args.cap = std::min(args.cap, SPREAD_THREADS / NDIM); // one stencil task per thread and window
        const size_t smem = sizeof(double) * ((size_t)RPTS + (size_t)args.cap * NDIM * W) + 2 * sizeof(int) * (size_t)args.cap;
        // L2 promotion of the block loads: 64 B measured best
        for (int a = 0; a < tp.ncomp; ++a)
            if (!no_tma && ((tp.comp[a].pp0[0] + M + XO) % 2 == 0) && make_tensor_map(&maps.m[a], tp.comp[a], NDIM, RX, R, R, 1))
                args.tma_mask |= (1u << a);
        auto kfn = spread_tile_kernel<NDIM, K>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
        {
            err = "cudaFuncSetAttribute(spread) failed";
            return e;
        }
        // 2^ndim tile colours, one launch each (same-colour blocks are disjoint)
        const int ncol = (NDIM == 3) ? 8 : 4;
        for (int c = 0; c < ncol; ++c)
        {
            int ntiles = 1;
            for (int d = 0; d < 3; ++d)
            {
                args.colour[d] = (d < NDIM) ? (c >> d) & 1 : 0;
                args.ntc[d] = (d < NDIM) ? (tp.nt[d] - args.colour[d] + 1) / 2 : 1;
                ntiles *= args.ntc[d];
            }
            if (ntiles <= 0) continue;
            kfn<<<dim3((unsigned)ntiles, (unsigned)tp.ncomp), SPREAD_THREADS, smem, L.stream>>>(tp, maps, args);
            L.launches++;
        }
    }
    ffn<<<1, FIXUP_THREADS, 0, L.stream>>>(tp, args);
    L.launches++;
    return cudaGetLastError();
}

template <int NDIM>
static cudaError_t launch_spread_k(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                                   std::string& err)
{
    switch (kernel)
    {
#define IBK_CASE(KK) \
    case KK:         \
        return launch_spread_t<NDIM, KK>(L, tp, bins, mv, err);
        IBK_CASE(IBK_PIECEWISE_LINEAR)
        IBK_CASE(IBK_IB_4)
        IBK_CASE(IBK_IB_6)
        IBK_CASE(IBK_BSPLINE_3)
        IBK_CASE(IBK_BSPLINE_4)
        IBK_CASE(IBK_IB_3)
        IBK_CASE(IBK_BSPLINE_5)
        IBK_CASE(IBK_BSPLINE_6)
        IBK_CASE(IBK_PIECEWISE_CUBIC)
        IBK_CASE(IBK_IB_5)
        IBK_CASE(IBK_PIECEWISE_CONSTANT)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_32)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_23)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_43)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_34)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_54)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_45)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_65)
        IBK_CASE(IBK_COMPOSITE_BSPLINE_56)
        IBK_CASE(IBK_DISCONTINUOUS_LINEAR)
        IBK_CASE(IBK_IB_4_W8)
#undef IBK_CASE
    default:
        err = "unknown kernel";
        return cudaErrorInvalidValue;
    }
}

cudaError_t launch_spread(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                          std::string& err)
{
    if (tp.ndim == 3) return launch_spread_k<3>(L, kernel, tp, bins, mv, err);
    return launch_spread_k<2>(L, kernel, tp, bins, mv, err);
}

} // namespace ibk
