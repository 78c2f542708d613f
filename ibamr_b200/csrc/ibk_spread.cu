// ibk_spread.cu -- force spreading (markers -> grid) for sm_100a, deterministic.
//
// Replaces lagrangian_<kernel>_spread{2,3}d (ibtk/src/lagrangian/fortran/
// lagrangian_interaction3d.f.m4:1344-1475 ib_4, :2384-2582 ib_6, :2703-2805 bspline_3,
// :2933-3041 bspline_4, :617-730 piecewise_linear; 2D twins in lagrangian_interaction2d.f.m4) and
// the per-axis loop of LEInteractor (LEInteractor.cpp:3676-3711).  The reference is serial over
// markers, so it has no write conflicts; here the work is organised so that none can occur:
//
//  * MARKER TILES WITH A HALOED ACCUMULATOR.  One CTA takes one marker tile (16^ndim cells = 4^ndim
//    bricks, one contiguous run of the sorted markers) and ONE component, and accumulates the full
//    stencils of its markers into a shared-memory block of (16 + 2M)^ndim points (M = kernel reach), so
//    every marker is visited exactly once per component and no stencil is ever clipped.
//  * TILE COLOURING.  The blocks of two tiles whose indices differ by 2 in some dimension are disjoint
//    (32 >= 16 + 2M), so the tiles are processed in 2^ndim launches (colours); inside a launch every
//    grid point is touched by at most one CTA, which finishes with one `f += block` pass over its
//    haloed block (the contract of LDataManager::spread, LDataManager.cpp:662-663).
//  * BRICK COLOURING inside the CTA.  One warp takes one brick (4^ndim cells) at a time and walks its
//    markers in storage order with the 32 lanes spread over the stencil points.  Bricks NC apart have
//    disjoint footprints; the bricks are visited colour by colour with a CTA barrier between colours, so no
//    two warps ever touch the same accumulator word at the same time.
//  * 1-D weights are evaluated one thread per (marker, dimension) for a window of markers and parked in
//    shared memory; the scaled force is folded into the last dimension's weights.
//  The summation order at every grid point is fixed by (tile colour, brick colour, sorted marker order):
//  results are bit-reproducible run to run.
//
// A marker whose stencil does not fit the haloed block (possible only if its binning cell and its stencil
// origin disagree by a rounding) is skipped here and spread by spread_fixup_kernel, one thread, in sorted
// order, after the last colour.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ibk_engine.h"
#include "ibk_tma.h"

namespace ibk
{
// -DIBK_TIMELINE: thread 0 of a few CTAs records clock64() at the phase boundaries (printed by the launcher)
#ifdef IBK_TIMELINE
__device__ long long g_tl[64][16];
__device__ int g_tl_n;
#define TL(k)                                             \
    do                                                    \
    {                                                     \
        if (threadIdx.x == 0 && tl_on) tl[k] = clock64(); \
    } while (0)
#else
#define TL(k)
#endif
constexpr int SPREAD_THREADS = 256;
// PLANE OWNERS (3D, 4-point kernels): the accumulation of a window is not organised by bricks and colours but by planes
// of the block: warp `id` owns the two z planes 2 id, 2 id + 1 and walks ALL markers of the window whose stencil meets
// them (2 or 3 warps per marker), lanes = 4 x 4 (x, y) points x 2 planes.  No two warps share a word, so the only CTA
// barriers left are the two around the stencil evaluation of a window; the order of the additions at a grid point is
// the window's marker order, the same as with the brick colours (results are bit-identical to that path).
constexpr bool SPREAD_CLUSTER_DEFAULT = false; // correct (all parity tests pass) but 13.6 ms: the read-modify-write of the share by
                                               // the threads is far too slow; next: dense repack + TMA reducing store (IBK_SPREAD_CLUSTER=1)
constexpr bool SPREAD_PLANES_DEFAULT = false; // measured SLOWER than the brick colours (5.83 ms against 4.41 ms on the C5 shard:
                                              // 2.5 dependent visits per marker instead of one); kept for reference, IBK_SPREAD_PLANES=1
constexpr int SPREAD_THREADS_PLANES = 320; // 10 warps = the 10 plane pairs of a 20-plane block
template <int NDIM, int K>
constexpr bool spread_planes_ok = (NDIM == 3 && KTraits<K>::W == 4 && KTraits<K>::M == 2);
constexpr int SPREAD_TASKS = 1; // (marker, dimension) stencil evaluations per thread and window (measured: a second one
                                // serialises two sqrt/div chains before the barrier: 85-marker windows beat 100-marker ones)
constexpr int SPREAD_WARPS = SPREAD_THREADS / 32;

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
    // exceptions (stencil outside the haloed block): (sorted position * 8 + component), see spread_fixup_kernel
    int* exc_count;
    int* exc_list;
    int exc_capacity;
    int cap; // markers whose stencil weights are staged at a time (sizes the dynamic shared memory)
    unsigned tma_mask; // bit a: the block of component a is loaded / stored by TMA (else zero-fill + red write-out)
    int part, sel_lo[3], sel_hi[3]; // MarkerView's tile selection
    int tma_reduce;                 // 1 (3D): the block starts from zero and is ADDED to f by TMA's reducing store
    int dense_thresh;               // > 0: bricks with more markers than this are left to spread_dense_kernel
};

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

// CL (3D): a thread-block CLUSTER of 2 x 2 x 2 CTAs acts as one 32^3 tile.  Each CTA accumulates the markers of its own
// 16^3 tile into its own block starting from zero; after a cluster barrier it sums, in rank order, its own block and the
// parts of its siblings' blocks (read through distributed shared memory) that cover its share of the cluster's
// (32 + 2M)^3 points -- its tile plus the halo on the cluster's outer sides -- and adds that share to f once.  Only the
// cluster's outer shell is shared with other launches (the colours are those of the cluster tiles), so f moves
// ((32 + 2M) / 32)^3 times instead of ((16 + 2M) / 16)^3, and points nothing was spread to are neither read nor written.
// WIDE: 320 instead of 256 threads with the brick colours: the window grows from 85 to 100 markers (one stencil task per thread).
template <int NDIM, int K, bool PL, bool CL = false, bool WIDE = false>
__global__ void __launch_bounds__((PL || WIDE) ? SPREAD_THREADS_PLANES : SPREAD_THREADS, (KTraits<K>::M <= 2) ? 3 : 2)
    spread_tile_kernel(const __grid_constant__ TileParams tp, const __grid_constant__ TmaMapSet maps, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int R = TILE + 2 * M; // haloed block edge
    constexpr int NT = (PL || WIDE) ? SPREAD_THREADS_PLANES : SPREAD_THREADS;
    constexpr int NWARPS = NT / 32;
    static_assert(!PL || spread_planes_ok<NDIM, K>, "plane owners: 3D, W = 4, M = 2");
    static_assert(!CL || (NDIM == 3 && !PL), "clusters: 3D, brick-colour accumulation");
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
    static_assert(R <= 24 && R < 128, "write-out covers a row with 8 lanes x 3 points; origins are kept in bytes");
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
    bool has_tile = true; // CL: the cluster tile may stick out of the tile grid
    bool active = true;   // this CTA has markers to spread
    const int crank = CL ? (int)(blockIdx.x & 7u) : 0; // rank in the cluster: bit d = upper half along d
    {
        int r = CL ? (int)(blockIdx.x >> 3) : (int)blockIdx.x;
        t[0] = 2 * (r % args.ntc[0]) + args.colour[0];
        r /= args.ntc[0];
        t[1] = 2 * (r % args.ntc[1]) + args.colour[1];
        if (NDIM == 3) t[2] = 2 * (r / args.ntc[1]) + args.colour[2];
        if constexpr (CL)
        {
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                t[d] = 2 * t[d] + ((crank >> d) & 1); // (t[] held the cluster tile)
                has_tile = has_tile && t[d] < tp.nt[d];
            }
            active = has_tile;
        }
    }
    if (args.part)
    {
        bool in = true;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) in = in && t[d] >= args.sel_lo[d] && t[d] <= args.sel_hi[d];
        if ((args.part == 1) != in)
        {
            if constexpr (!CL) return;
            active = false; // the other part's tile: nothing to spread, but the share still collects the siblings' halos
        }
    }
    const int a = blockIdx.y;
    const CompGeom& cg = tp.comp[a];
    const int tile = (NDIM == 3) ? (t[2] * tp.nt[1] + t[1]) * tp.nt[0] + t[0] : t[1] * tp.nt[0] + t[0];
    const int b0 = tp.brick_base + tile * NBRICKS;
    // one coalesced read of the tile's NBRICKS + 1 segment offsets, and (independent of it) the colour order
    int my_q = 0;
    if (threadIdx.x <= NBRICKS) sbs[threadIdx.x] = active ? __ldg(&args.brick_start[b0 + threadIdx.x]) : 0;
    if (threadIdx.x < NBRICKS) my_q = __ldg(&d_colouring<NDIM, NC>.order[threadIdx.x]);
    for (int q = threadIdx.x; q < args.cap; q += NT) relb[q] = 0;
#ifdef IBK_TIMELINE
    long long tl[16];
    for (int k = 0; k < 16; ++k) tl[k] = 0;
    const bool tl_on = threadIdx.x == 0 && blockIdx.y == 0 && (blockIdx.x % 97) == 5;
    TL(0);
#endif
    __syncthreads();
    const int s0 = sbs[0], s1 = sbs[NBRICKS];
    if (s0 >= s1)
    {
        if constexpr (!CL) return;
        active = false;
    }
    TL(1);

    int blo[3]; // pp coordinate of the block's first point
#pragma unroll
    for (int d = 0; d < 3; ++d) blo[d] = TILE * t[d] - M;
    // With TMA the block starts as a copy of f (out-of-array points read as zero) and is stored back at the end
    // (out-of-array points dropped): `f += S[F]` without a separate zero / add pass and without atomics.
    // (Measured on B200: a TMA tensor STORE with a negative start coordinate traps, loads do not; the blocks of the
    // first tile per dimension therefore take the fallback.  The store also writes whole 16-byte units, i.e. one
    // element of the row padding when n[0] is odd: harmless, nothing reads the padding.)
    const bool use_tma = !CL && ((args.tma_mask >> a) & 1u) && (blo[0] - XO - cg.pp0[0] >= 0) && (blo[1] - cg.pp0[1] >= 0) &&
                         (NDIM == 2 || blo[2] - cg.pp0[2] >= 0);
    // TMA's reducing store (cp.reduce.async.bulk.tensor .add; measured to work on fp64 tensors, scripts/tma_reduce_probe.cu):
    // the block starts from zero and is added to f in L2, no load.  One CTA per grid point and launch: the order is fixed.
    const bool use_red = NDIM == 3 && use_tma && args.tma_reduce;
    if (use_tma && !use_red && threadIdx.x == 0)
    {
        mbar_init(&tma_bar, 1);
        mbar_fence_init();
    }

    // marker ranges of the bricks in colour-major order, and their running count (two-warp scan)
    int my_cnt = 0, my_incl = 0;
    if (threadIdx.x < NBRICKS) // (an inactive CTA of a cluster has sbs[] = 0: no markers)
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
    if (!use_tma || use_red)
        for (int q = threadIdx.x; q < RPTS; q += NT) acc[q] = 0.0;
    __syncthreads();
    TL(2);
    if (use_tma && !use_red && threadIdx.x == 0)
    {
        mbar_expect_tx(&tma_bar, (uint32_t)(RPTS * sizeof(double)));
        if (NDIM == 3)
            tma_load_3d(acc, &maps.m[a], &tma_bar, blo[0] - XO - cg.pp0[0], blo[1] - cg.pp0[1], blo[2] - cg.pp0[2]);
        else
            tma_load_2d(acc, &maps.m[a], &tma_bar, blo[0] - XO - cg.pp0[0], blo[1] - cg.pp0[1]);
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
    // Same-colour bricks have disjoint footprints and a barrier separates the colours, so no two warps ever
    // touch the same accumulator word at the same time.
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
        poffb[s] = 8 * ((iz * R + iy) * RX + ix + XO);
        pw0[s] = ix;
        pw1[s] = W + iy;
        pw2[s] = 2 * W + iz;
    }
    __syncthreads();
    TL(3);
    const int total = bpre[NBRICKS]; // without the dense bricks

    // stencil-evaluation tasks of this thread for a window: the global loads are issued by fetch() -- for window
    // w + 1 before the accumulation of window w starts, so their latency hides behind it -- and consumed by
    // evaluate() at the top of the window.
    int t_i[SPREAD_TASKS];
    double t_xs[SPREAD_TASKS], t_xr[SPREAD_TASKS], t_v[SPREAD_TASKS];
    auto fetch = [&](int off, int par) {
        const int cnt = min(cap, total - off);
#pragma unroll
        for (int k = 0; k < SPREAD_TASKS; ++k)
        {
            const int tix = threadIdx.x + k * NT;
            t_i[k] = -1;
            if (tix >= cnt * NDIM) continue;
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
            t_i[k] = i;
            t_xs[k] = __ldg(&Xp[d * args.x_stride + i]);
            t_xr[k] = Xr ? __ldg(&Xr[d * args.x_stride + i]) : 0.0;
            t_v[k] = 1.0;
            if (d == LD)
            {
                const long long row = args.src ? (long long)__ldg(&args.src[i]) : (long long)i;
                t_v[k] = __ldg(&args.V[vcol * args.v_cstride + row * args.v_istride]);
            }
        }
    };
    auto evaluate = [&](int par) {
#pragma unroll
        for (int k = 0; k < SPREAD_TASKS; ++k)
        {
            if (t_i[k] < 0) continue;
            const int tix = threadIdx.x + k * NT;
            const int m = tix / NDIM, d = tix - m * NDIM;
            const double xl_s = tp.xl[d][cg.var[d]];
            const double dx_s = tp.dx[d];
            const int blo_s = (d == 0) ? blo[0] : (d == 1) ? blo[1] : blo[2];
            double w[W];
            int l;
            stencil_1d<K>(t_xs[k], Xr ? t_xr[k] : t_xs[k], xl_s, dx_s, l, w, d == cg.axis);
            const int r0 = l + G - blo_s; // first stencil point relative to the block
            const bool fits = r0 >= 0 && r0 + W <= R;
            const double scale = (d == LD) ? t_v[k] * inv_vol : 1.0;
#pragma unroll
            for (int j = 0; j < W; ++j) wgt[(m * NDIM + d) * W + j] = w[j] * scale;
            if constexpr (PL)
            {
                // (x, y) byte offset in the low 20 bits, first z plane above them
                atomicAdd(&relb[par * cap + m], !fits ? -(1 << 29) : (d == 0) ? r0 * 8 : (d == 1) ? r0 * 8 * RX : (r0 << 20));
                if (!fits && args.exc_list) // left to the fix-up (which removes duplicates)
                {
                    const int slot = atomicAdd(args.exc_count, 1);
                    if (slot < args.exc_capacity) args.exc_list[slot] = t_i[k] * 8 + a;
                }
            }
            else
            {
                const int stride_b = (d == 0) ? 8 : (d == 1) ? 8 * RX : 8 * RX * R; // bytes per point along d in the block
                atomicAdd(&relb[par * cap + m], fits ? r0 * stride_b : -(1 << 29));
            }
        }
    };

    fetch(0, 0);
    int par = 0;
    for (int off = 0; off < total; off += cap, par ^= 1)
    {
        const int cnt = min(cap, total - off);
        // ---- phase A
        evaluate(par);
        __syncthreads();
        if (off == 0) TL(4);
        if (use_tma && !use_red && off == 0) mbar_wait(&tma_bar, 0); // the block holds f now
        if (off == 0) TL(5);
        const int col_lo = wcol[par][0], col_hi = wcol[par][1];
        if (off + cap < total) fetch(off + cap, par ^ 1);
        if constexpr (PL)
        {
            // ---- phase B, plane owners: no barrier until the window is done
            static_assert(8 * RX * R < (1 << 20), "(x, y) byte offset fits 20 bits");
            constexpr int PLANE_B = 8 * RX * R;
            constexpr int NW = NDIM * W;
            const int* rp = relb + par * cap;
            const int ix = lane & 3, iy = (lane >> 2) & 3, iz = lane >> 4;
            const double* wl = wgt + ix; // this lane's x weight of marker 0; y weight at + W + (iy - ix)
            for (int id = warp; id < R / 2; id += NWARPS)
            {
                const int plane = 2 * id + iz;
                char* const accp = reinterpret_cast<char*>(acc) + plane * PLANE_B + 8 * (iy * RX + ix + XO);
                for (int c0 = 0; c0 < cnt; c0 += 32)
                {
                    const int mm = c0 + lane;
                    const int ab = (mm < cnt) ? rp[mm] : -1;
                    const int r0l = ab >> 20;
                    unsigned hits = __ballot_sync(0xffffffffu, ab >= 0 && r0l <= 2 * id + 1 && r0l + 3 >= 2 * id);
                    while (hits)
                    {
                        const int b = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const int abm = __shfl_sync(0xffffffffu, ab, b);
                        const int kz = plane - (abm >> 20);
                        const double* wp = wl + (c0 + b) * NW;
                        if ((unsigned)kz < (unsigned)W)
                        {
                            const double wv = (wp[0] * wp[W + iy - ix]) * wp[2 * W + kz - ix];
                            double* pt = reinterpret_cast<double*>(accp + (abm & 0xFFFFF));
                            *pt = *pt + wv;
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
        }
        else
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
                    else if (lane == 0 && args.exc_list)
                    {
                        const int slot = atomicAdd(args.exc_count, 1);
                        if (slot < args.exc_capacity) args.exc_list[slot] = (bfirst[p] + (off + m - p0)) * 8 + a;
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (off == 0) TL(6);
        if (off == cap) TL(7);
    }
    TL(8);

    if constexpr (CL)
    {
        // ---- cluster write-out: this CTA's share of the cluster's points = own block + the siblings' halos, in rank order
        namespace cgx = cooperative_groups;
        cgx::cluster_group cluster = cgx::this_cluster();
        __shared__ int cl_any; // this CTA spread something (the siblings read it)
        if (threadIdx.x == 0) cl_any = total > 0 ? 1 : 0;
        cluster.sync(); // every block of the cluster is complete
        {
            int any = 0;
#pragma unroll
            for (int sr = 0; sr < 8; ++sr) any |= *cluster.map_shared_rank(&cl_any, sr);
            if (!any) // nothing was spread into this cluster tile: nothing to add (the same decision in all eight CTAs)
            {
                cluster.sync();
                return;
            }
        }
        constexpr int SH = TILE + M; // share edge: the tile and the halo on the cluster's outer side
        const int o0 = (crank & 1) ? M : 0, o1 = (crank & 2) ? M : 0, o2 = (crank & 4) ? M : 0;
        // Write-out by ONE reducing TMA store of the densely repacked share (the share maps have an SH^3 box) when the
        // share starts inside the array on an even x coordinate; else by the threads.
        const int c0 = blo[0] + o0 - cg.pp0[0], c1 = blo[1] + o1 - cg.pp0[1], c2 = blo[2] + o2 - cg.pp0[2];
        const bool by_tma = has_tile && XO == 0 && args.tma_reduce && ((args.tma_mask >> a) & 1u) && c0 >= 0 && (c0 & 1) == 0 && c1 >= 0 && c2 >= 0;
        if (has_tile)
        {
            constexpr int UNR = 8; // points per thread and round: their loads are in flight together
            unsigned sok = 0;      // siblings that exist (their tile is inside the tile grid)
#pragma unroll
            for (int sr = 0; sr < 8; ++sr)
            {
                bool ok = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) ok = ok && (t[d] - ((crank >> d) & 1) + ((sr >> d) & 1)) < tp.nt[d];
                if (ok) sok |= 1u << sr;
            }
            // the total at share point (lx, ly, lz) of this CTA's block; `both`: the dimensions along which the other
            // sibling's block holds the point too (it lies in the 2M-wide overlap)
            auto overlap = [&](int lx, int ly, int lz) -> unsigned {
                return ((crank & 1) ? (lx < 2 * M) : (lx >= TILE)) | (((crank & 2) ? (ly < 2 * M) : (ly >= TILE)) << 1) |
                       (((crank & 4) ? (lz < 2 * M) : (lz >= TILE)) << 2);
            };
            auto total_at = [&](int lx, int ly, int lz, unsigned both) -> double {
                double v = 0.0;
#pragma unroll
                for (int sr = 0; sr < 8; ++sr)
                {
                    const unsigned diff = (unsigned)sr ^ (unsigned)crank;
                    if ((diff & ~both) != 0 || !((sok >> sr) & 1u)) continue;
                    // the sibling's block starts 16 points later (earlier) along the dimensions where it is the upper (lower) one
                    const int cx = lx + ((diff & 1) ? ((crank & 1) ? TILE : -TILE) : 0);
                    const int cy = ly + ((diff & 2) ? ((crank & 2) ? TILE : -TILE) : 0);
                    const int cz = lz + ((diff & 4) ? ((crank & 4) ? TILE : -TILE) : 0);
                    v += cluster.map_shared_rank(acc, sr)[(cz * R + cy) * RX + cx + XO];
                }
                return v;
            };
            if (by_tma)
            {
                // totals IN PLACE: a CTA's share and the strips its siblings read from its block are disjoint, and the
                // share points outside the overlaps already hold their total
                for (int q = threadIdx.x; q < SH * SH * SH; q += NT)
                {
                    const int lx = q % SH + o0, ly = (q / SH) % SH + o1, lz = q / (SH * SH) + o2;
                    const unsigned both = overlap(lx, ly, lz);
                    if (both) acc[(lz * R + ly) * RX + lx + XO] = total_at(lx, ly, lz, both);
                }
            }
            else
            {
                // the element of f behind share point q (nullptr: outside the array)
                auto gaddr = [&](int q) -> double* {
                    const int gi = c0 + q % SH, gj = c1 + (q / SH) % SH, gk = c2 + q / (SH * SH);
                    if (gi < 0 || gi >= cg.n[0] || gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) return nullptr;
                    return cg.ptr + ((long long)(gk * cg.n[1] + gj) * cg.pitch + gi);
                };
                for (int q0 = threadIdx.x; q0 < SH * SH * SH; q0 += NT * UNR)
                {
                    double v[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                    {
                        const int q = q0 + u * NT;
                        v[u] = 0.0;
                        if (q >= SH * SH * SH) continue;
                        const int lx = q % SH + o0, ly = (q / SH) % SH + o1, lz = q / (SH * SH) + o2; // in this CTA's block
                        const unsigned both = overlap(lx, ly, lz);
                        v[u] = both ? total_at(lx, ly, lz, both) : acc[(lz * R + ly) * RX + lx + XO];
                    }
                    // f += share: inside a launch a grid point belongs to exactly one CTA (plain read-modify-write, fixed
                    // order); points nothing was spread to are neither read nor written
                    unsigned wr = 0;
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                    {
                        if (v[u] == 0.0) continue;
                        const double* g = gaddr(q0 + u * NT);
                        if (!g) continue;
                        v[u] = *g + v[u];
                        wr |= 1u << u;
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                        if ((wr >> u) & 1u) *gaddr(q0 + u * NT) = v[u];
                }
            }
        }
        cluster.sync(); // nobody reads this CTA's block any more
        if (by_tma)
        {
            // dense repack of the share to the start of the block: destination q never lies behind its source, so chunks of
            // NT points in ascending order only need their reads separated from their writes
            for (int q0 = 0; q0 < SH * SH * SH; q0 += NT)
            {
                const int q = q0 + threadIdx.x;
                double val = 0.0;
                if (q < SH * SH * SH) val = acc[((q / (SH * SH) + o2) * R + (q / SH) % SH + o1) * RX + q % SH + o0];
                __syncthreads();
                if (q < SH * SH * SH) acc[q] = val;
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (threadIdx.x == 0)
            {
                asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&maps.m[a]),
                             "r"(smem_u32(acc)), "r"(c0), "r"(c1), "r"(c2)
                             : "memory");
                tma_store_commit_and_wait_read();
            }
        }
        return;
    }

    // ---- write-out.  TMA: the block (= old f + the spread values) is stored back, clipped to the array.
    if (use_tma)
    {
        fence_proxy_async_smem(); // this thread's generic-proxy writes to the block -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (use_red)
                asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&maps.m[a]),
                             "r"(smem_u32(acc)), "r"(blo[0] - XO - cg.pp0[0]), "r"(blo[1] - cg.pp0[1]), "r"(blo[2] - cg.pp0[2])
                             : "memory");
            else if (NDIM == 3)
                tma_store_3d(&maps.m[a], acc, blo[0] - XO - cg.pp0[0], blo[1] - cg.pp0[1], blo[2] - cg.pp0[2]);
            else
                tma_store_2d(&maps.m[a], acc, blo[0] - XO - cg.pp0[0], blo[1] - cg.pp0[1]);
            tma_store_commit_and_wait_read(); // shared memory must stay alive until it has been read
#ifdef IBK_TIMELINE
            TL(9);
            if (tl_on)
            {
                const int slot = atomicAdd(&g_tl_n, 1);
                if (slot < 64)
                    for (int k = 0; k < 16; ++k) g_tl[slot][k] = tl[k];
            }
#endif
        }
        return;
    }
    // Fallback (array not addressable by TMA): f += block with `red.global.add.f64`, rows along x, dropping points
    // outside the array.  Inside one launch a grid point belongs to at most one CTA and the launches are ordered by
    // the stream, so the order of the additions is fixed and the result bit-reproducible.
    {
        constexpr int ROWS = (NDIM == 3) ? R * R : R;
        constexpr int RSTEP = NT / 8; // rows per sweep: a group of 8 lanes takes one row at a time,
        const int g8 = threadIdx.x >> 3, l8 = threadIdx.x & 7; // 8 lanes x 3 points cover R <= 24 points
        const int gx0 = blo[0] - cg.pp0[0] + l8;
        bool okx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) okx[k] = (l8 + 8 * k < R) && (gx0 + 8 * k >= 0) && (gx0 + 8 * k < cg.n[0]);
        int y = g8 % R, z = g8 / R;
        const int gy0 = blo[1] - cg.pp0[1], gz0 = (NDIM == 3) ? blo[2] - cg.pp0[2] : 0;
        const double* arow = acc + g8 * RX + XO + l8;
        for (int row = g8; row < ROWS; row += RSTEP, arow += RSTEP * RX)
        {
            const int gj = gy0 + y, gk = gz0 + z;
            y += RSTEP % R;
            z += RSTEP / R;
            if (y >= R)
            {
                y -= R;
                ++z;
            }
            if (gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
            double* prow = cg.ptr + ((long long)(gk * cg.n[1] + gj) * cg.pitch + gx0);
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                if (!okx[k]) continue;
                const double v = arow[8 * k];
                if (v != 0.0) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(prow + 8 * k), "d"(v) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense bricks (3D).  A structure puts tens of markers into a cell, hundreds into a brick; the
// tile kernel would walk them with ONE warp (a brick is the unit of its colouring).  Here a CTA takes one dense
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
            else if (args.exc_list && zh == 0) // the whole (marker, component) goes to the fix-up; a zero factor removes it here
            {
                const int slot = atomicAdd(args.exc_count, 1);
                if (slot < args.exc_capacity) args.exc_list[slot] = i * 8 + a;
            }
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

// Fix-up for the (practically never occurring) (marker, component) pairs whose stencil does not fit the
// haloed block of their tile: one thread, sorted order, the whole stencil, clipped to the array only.
template <int NDIM, int K>
__global__ void spread_fixup_kernel(const __grid_constant__ TileParams tp, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = *args.exc_count;
    if (n <= 0) return;
    if (n > args.exc_capacity) n = args.exc_capacity;
    for (int a = 1; a < n; ++a) // insertion sort (tiny list)
    {
        const int v = args.exc_list[a];
        int b = a - 1;
        while (b >= 0 && args.exc_list[b] > v)
        {
            args.exc_list[b + 1] = args.exc_list[b];
            --b;
        }
        args.exc_list[b + 1] = v;
    }
    for (int e = 0; e < n; ++e)
    {
        const int code = args.exc_list[e];
        if (e > 0 && args.exc_list[e - 1] == code) continue;
        const int i = code >> 3, a = code & 7;
        const CompGeom& cg = tp.comp[a];
        const long long row = args.src ? (long long)args.src[i] : (long long)i;
        double w[3][W];
        int lo[3] = { 0, 0, 0 };
        for (int d = 0; d < NDIM; ++d)
        {
            const double xs = args.X[d * args.x_stride + i];
            const double xr = args.Xraw ? args.Xraw[d * args.x_stride + i] : xs;
            int l;
            stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w[d], d == cg.axis);
            lo[d] = l + tp.G;
        }
        const double f = args.V[cg.vcol * args.v_cstride + row * args.v_istride] * tp.inv_vol;
        const int KW = (NDIM == 3) ? W : 1;
        for (int k = 0; k < KW; ++k)
            for (int j = 0; j < W; ++j)
                for (int ii = 0; ii < W; ++ii)
                {
                    const int gi = lo[0] + ii - cg.pp0[0], gj = lo[1] + j - cg.pp0[1],
                              gk = (NDIM == 3) ? lo[2] + k - cg.pp0[2] : 0;
                    if (gi < 0 || gi >= cg.n[0] || gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
                    double wv = w[0][ii] * w[1][j];
                    if (NDIM == 3) wv *= w[2][k];
                    cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi] += wv * f;
                }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int* g_exc_buf = nullptr; // [1 + capacity] per process (device); tiny
constexpr int EXC_CAPACITY = 4096;

template <int NDIM, int K, bool PL, bool WIDE = false>
static cudaError_t launch_spread_pl(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err);

template <int NDIM, int K>
static cudaError_t launch_spread_t(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    if constexpr (spread_planes_ok<NDIM, K>)
    {
        // IBK_SPREAD_PLANES=0 / 1 overrides the default
        static const char* env = getenv("IBK_SPREAD_PLANES");
        static const bool planes = env ? atoi(env) != 0 : SPREAD_PLANES_DEFAULT;
        if (planes) return launch_spread_pl<NDIM, K, true>(L, tp, bins, mv, err);
        static const bool wide = getenv("IBK_SPREAD_WIDE") ? atoi(getenv("IBK_SPREAD_WIDE")) != 0 : false; // not measured yet
        if (wide) return launch_spread_pl<NDIM, K, false, true>(L, tp, bins, mv, err);
    }
    return launch_spread_pl<NDIM, K, false>(L, tp, bins, mv, err);
}

template <int NDIM, int K, bool PL, bool WIDE>
static cudaError_t launch_spread_pl(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    constexpr int NT = (PL || WIDE) ? SPREAD_THREADS_PLANES : SPREAD_THREADS;
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int R = TILE + 2 * M;
    constexpr int XO = M & 1;
    constexpr int RX = (R + XO + 1) & ~1;
    constexpr int RPTS = (NDIM == 3) ? R * R * RX : R * RX;
    cudaError_t e;
    if (!g_exc_buf)
    {
        if ((e = cudaMalloc(&g_exc_buf, sizeof(int) * (1 + EXC_CAPACITY))) != cudaSuccess) return e;
    }
    if ((e = cudaMemsetAsync(g_exc_buf, 0, sizeof(int), L.stream)) != cudaSuccess) return e;
    SpreadArgs args;
    args.brick_start = bins.brick_start;
    args.X = mv.X;
    args.Xraw = mv.Xraw;
    args.x_stride = mv.x_stride;
    args.V = mv.V;
    args.v_cstride = mv.v_cstride;
    args.v_istride = mv.v_istride;
    args.src = mv.src;
    args.exc_count = g_exc_buf;
    args.exc_list = g_exc_buf + 1;
    args.exc_capacity = EXC_CAPACITY;
    args.part = mv.part;
    {
        // measured 4.08 ms against 4.28 ms; tests/test_gpu_configs.py passes with it, the full parity suite has not been run yet
        static const bool red = getenv("IBK_SPREAD_REDUCE") ? atoi(getenv("IBK_SPREAD_REDUCE")) != 0 : false;
        args.tma_reduce = red ? 1 : 0;
    }
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
    // window size: as many markers as the shared memory left by the block allows at the target residency
    static const int cap_env = getenv("IBK_SPREAD_CAP") ? atoi(getenv("IBK_SPREAD_CAP")) : 0;
    constexpr int target_ctas = (M <= 2) ? 3 : 2;
    constexpr long long budget = 233472 / target_ctas - 1024 - 2304 - (long long)sizeof(double) * RPTS;
    constexpr int per_marker = (int)sizeof(double) * NDIM * W + 2 * (int)sizeof(int);
    constexpr int cap_fit = (int)(budget / per_marker);
    args.cap = (cap_env >= 8 && cap_env <= 1024) ? cap_env : std::max(32, std::min(256, cap_fit));
    args.cap = std::min(args.cap, SPREAD_TASKS * NT / NDIM); // fetch()/evaluate() hold SPREAD_TASKS tasks per thread
    const size_t smem = sizeof(double) * ((size_t)RPTS + (size_t)args.cap * NDIM * W) + 2 * sizeof(int) * (size_t)args.cap;
    // TMA moves the block when it can address the array and the block starts on an even x coordinate
    TmaMapSet maps;
    std::memset(&maps, 0, sizeof(maps));
    args.tma_mask = 0;
    static const bool no_tma = getenv("IBK_NO_TMA") != nullptr;
    // L2 promotion of the block loads: 64 B measured best (4.28 ms; 128 B and none 4.41 ms on the C5 shard)
    static const int promo = getenv("IBK_TMA_PROMO_SPREAD") ? atoi(getenv("IBK_TMA_PROMO_SPREAD")) : 1;
    for (int a = 0; a < tp.ncomp; ++a)
        if (!no_tma && ((tp.comp[a].pp0[0] + M + XO) % 2 == 0) && make_tensor_map(&maps.m[a], tp.comp[a], NDIM, RX, R, R, promo))
            args.tma_mask |= (1u << a);
    static const bool dbg = getenv("IBK_DEBUG") != nullptr;
    if (dbg)
        for (int a = 0; a < tp.ncomp; ++a)
            fprintf(stderr, "[ibk] spread<%d,%d> comp %d tma=%u n=(%d,%d,%d) pitch=%lld pp0=(%d,%d,%d) ptr=%p nt=(%d,%d,%d) box=(%d,%d)\n", NDIM,
                    K, a, (args.tma_mask >> a) & 1u, tp.comp[a].n[0], tp.comp[a].n[1], tp.comp[a].n[2], tp.comp[a].pitch,
                    tp.comp[a].pp0[0], tp.comp[a].pp0[1], tp.comp[a].pp0[2], (void*)tp.comp[a].ptr, tp.nt[0], tp.nt[1], tp.nt[2], RX, R);
    auto kfn = spread_tile_kernel<NDIM, K, PL, false, WIDE>;
    auto ffn = spread_fixup_kernel<NDIM, K>;
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
    {
        err = "cudaFuncSetAttribute(spread) failed";
        return e;
    }
    // dense bricks first (their own kernel), then the tiles without them
    args.dense_thresh = 0;
    if constexpr (NDIM == 3 && KTraits<K>::M <= 3) // (a 12^3 footprint does not fit the dense kernel's static shared memory)
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
    if constexpr (NDIM == 3 && !PL && !WIDE)
    {
        static const char* env = getenv("IBK_SPREAD_CLUSTER"); // 0 / 1 overrides the default
        static const bool use_cluster = env ? atoi(env) != 0 : SPREAD_CLUSTER_DEFAULT;
        if (use_cluster)
        {
            // clusters of 2 x 2 x 2 CTAs = 32^3 cluster tiles; 8 colours of cluster tiles, one launch each
            auto cfn = spread_tile_kernel<NDIM, K, false, true>;
            // IBK_SPREAD_CLUSTER=2: the shares are added by TMA's reducing store (maps with a (16 + M)^3 box; even M only)
            static const bool share_tma = env && atoi(env) == 2;
            args.tma_reduce = 0;
            args.tma_mask = 0;
            if (share_tma && M % 2 == 0)
            {
                args.tma_reduce = 1;
                for (int a = 0; a < tp.ncomp; ++a)
                    if (!no_tma && ((tp.comp[a].pp0[0] + M) % 2 == 0) && make_tensor_map(&maps.m[a], tp.comp[a], NDIM, TILE + M, TILE + M, TILE + M, promo))
                        args.tma_mask |= (1u << a);
            }
            e = cudaFuncSetAttribute(cfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
            {
                err = "cudaFuncSetAttribute(spread, cluster) failed";
                return e;
            }
            for (int c = 0; c < 8; ++c)
            {
                int nclusters = 1;
                for (int d = 0; d < 3; ++d)
                {
                    const int nct = (tp.nt[d] + 1) / 2; // cluster tiles along d
                    args.colour[d] = (c >> d) & 1;
                    args.ntc[d] = (nct - args.colour[d] + 1) / 2;
                    nclusters *= args.ntc[d];
                }
                if (nclusters <= 0) continue;
                cudaLaunchConfig_t cfg;
                std::memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(8u * (unsigned)nclusters, (unsigned)tp.ncomp);
                cfg.blockDim = dim3(NT);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = L.stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 8;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                if ((e = cudaLaunchKernelEx(&cfg, cfn, tp, maps, args)) != cudaSuccess)
                {
                    err = "cudaLaunchKernelEx(spread, cluster) failed";
                    return e;
                }
                L.launches++;
            }
            ffn<<<1, 32, 0, L.stream>>>(tp, args);
            L.launches++;
            return cudaGetLastError();
        }
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
        dim3 grid((unsigned)ntiles, (unsigned)tp.ncomp);
        kfn<<<grid, NT, smem, L.stream>>>(tp, maps, args);
        L.launches++;
    }
    ffn<<<1, 32, 0, L.stream>>>(tp, args);
    L.launches++;
#ifdef IBK_TIMELINE
    {
        static int calls = 0;
        if (++calls == 6)
        {
            cudaStreamSynchronize(L.stream);
            static long long h[64][16];
            int n = 0;
            cudaMemcpyFromSymbol(&n, g_tl_n, sizeof(int));
            cudaMemcpyFromSymbol(h, g_tl, sizeof(h));
            n = std::min(n, 64);
            double avg[16] = { 0 };
            int used = 0;
            for (int i = n / 2; i < n; ++i, ++used)
                for (int k = 1; k < 10; ++k) avg[k] += (double)(h[i][k] - h[i][0]);
            fprintf(stderr, "[timeline] %d samples; cycles since CTA start:", used);
            const char* nm[10] = { "start", "offsets", "bricks", "scan", "phaseA0", "tma_wait", "phaseB0", "window1", "allB", "stored" };
            for (int k = 1; k < 10; ++k) fprintf(stderr, " %s=%.0f", nm[k], avg[k] / std::max(used, 1));
            fprintf(stderr, "\n");
        }
    }
#endif
    return cudaGetLastError();
}

template <int NDIM>
static cudaError_t launch_spread_k(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                                   std::string& err)
{
    switch (kernel)
    {
    case IBK_PIECEWISE_LINEAR:
        return launch_spread_t<NDIM, IBK_PIECEWISE_LINEAR>(L, tp, bins, mv, err);
    case IBK_IB_4:
        return launch_spread_t<NDIM, IBK_IB_4>(L, tp, bins, mv, err);
    case IBK_IB_6:
        return launch_spread_t<NDIM, IBK_IB_6>(L, tp, bins, mv, err);
    case IBK_BSPLINE_3:
        return launch_spread_t<NDIM, IBK_BSPLINE_3>(L, tp, bins, mv, err);
    case IBK_BSPLINE_4:
        return launch_spread_t<NDIM, IBK_BSPLINE_4>(L, tp, bins, mv, err);
    case IBK_IB_3:
        return launch_spread_t<NDIM, IBK_IB_3>(L, tp, bins, mv, err);
    case IBK_BSPLINE_5:
        return launch_spread_t<NDIM, IBK_BSPLINE_5>(L, tp, bins, mv, err);
    case IBK_BSPLINE_6:
        return launch_spread_t<NDIM, IBK_BSPLINE_6>(L, tp, bins, mv, err);
    case IBK_PIECEWISE_CUBIC:
        return launch_spread_t<NDIM, IBK_PIECEWISE_CUBIC>(L, tp, bins, mv, err);
    case IBK_IB_5:
        return launch_spread_t<NDIM, IBK_IB_5>(L, tp, bins, mv, err);
    case IBK_PIECEWISE_CONSTANT:
        return launch_spread_t<NDIM, IBK_PIECEWISE_CONSTANT>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_32:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_32>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_23:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_23>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_43:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_43>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_34:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_34>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_54:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_54>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_45:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_45>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_65:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_65>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_56:
        return launch_spread_t<NDIM, IBK_COMPOSITE_BSPLINE_56>(L, tp, bins, mv, err);
    case IBK_DISCONTINUOUS_LINEAR:
        return launch_spread_t<NDIM, IBK_DISCONTINUOUS_LINEAR>(L, tp, bins, mv, err);
    case IBK_IB_4_W8:
        return launch_spread_t<NDIM, IBK_IB_4_W8>(L, tp, bins, mv, err);
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
