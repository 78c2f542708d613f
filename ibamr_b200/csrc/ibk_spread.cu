// ibk_spread.cu -- force spreading (markers -> grid) for sm_100a, deterministic, no
// floating-point atomics anywhere.
//
// Replaces lagrangian_<kernel>_spread{2,3}d (ibtk/src/lagrangian/fortran/
// lagrangian_interaction3d.f.m4:1344-1475 ib_4, :2384-2582 ib_6, :2703-2805 bspline_3,
// :2933-3041 bspline_4, :617-730 piecewise_linear; 2D twins in lagrangian_interaction2d.f.m4) and
// the per-axis loop of LEInteractor (LEInteractor.cpp:3676-3711).  The reference is serial over
// markers, so it has no write conflicts; here the work is organised so that none can occur:
//
//  * STENCIL RECORDS (spread_records_kernel).  One thread per (marker, dimension, x_lower variant)
//    evaluates the stencil origin and the 1-D weights once per marker and stores them, with the
//    scaled force, as one 16-byte-aligned record per marker in sorted-marker order.  The sqrt/div
//    chains run at full occupancy here instead of inside the latency-critical tile kernel.
//  * OWNER-COMPUTES TILES (spread_tile_kernel).  One CTA owns a 16^ndim block of grid points of
//    every component and is the only writer of those points: it accumulates in shared memory and
//    finishes with one coalesced `f += tile` pass (the contract of LDataManager::spread,
//    LDataManager.cpp:662-663).  The CTA visits every marker whose stencil can reach its points: the
//    markers binned in the (16 + 2M)^ndim cells around the tile, i.e. NBR^ndim bricks of 4^ndim
//    cells, each a contiguous run of records.  Stencils are clipped to the tile.
//  * BRICK COLOURING.  Inside the CTA one warp takes one brick at a time and walks its markers in
//    storage order; the 32 lanes cover the stencil points.  Two bricks whose index differs by a
//    multiple of NC in every dimension have disjoint footprints (4*NC >= 4 + 2M), so the CTA runs
//    NC^ndim phases separated by __syncthreads and inside a phase no two warps touch the same
//    shared-memory word.  The summation order at every grid point is therefore fixed by the
//    sorted marker order alone: results are bit-reproducible run to run.
//
// Contributions to points farther than M cells from the marker's binning cell (possible only if
// binning cell and stencil origin disagree by a rounding) are excluded here by the margin mask and
// added by spread_fixup_kernel in a fixed order.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "ibk_engine.h"

namespace ibk
{
constexpr int SPREAD_THREADS = 256;
constexpr int SPREAD_WARPS = SPREAD_THREADS / 32;
constexpr int SPREAD_BATCH = 4; // markers per batch (per warp)
constexpr int SPREAD_MAXC = 3;  // components accumulated per launch
constexpr int REC_INTS = 10;    // lo[d][v] (6), binning cell cc[d] (3), spare

// doubles per marker record: weights [d < NDIM-1][v][j], last-dim weights * scaled force per component
// [a][j], integer section
template <int NDIM, int W>
struct RecLayout
{
    static constexpr int WGT = (NDIM - 1) * 2 * W;
    static constexpr int WLF = WGT;                      // offset of wlf[a][j]
    static constexpr int INTS = WGT + SPREAD_MAXC * W;   // offset (in doubles) of the integer section
    static constexpr int DOUBLES = ((INTS + REC_INTS / 2) + 1) / 2 * 2;
};

struct SpreadArgs
{
    const int* brick_start;
    const uint64_t* keys; // sorted keys (brick id, cell-in-brick bits)
    int tie_bits;
    const double* X;
    const double* Xraw;
    long long x_stride;
    const double* V;
    long long v_cstride, v_istride;
    const uint32_t* src;
    int comp0; // first component handled by this launch
    int ncomp; // number of components handled by this launch (<= SPREAD_MAXC)
    double* records; // [n_entries][RecLayout::DOUBLES]
    int first, last; // sorted positions of this patch's markers
    // exceptions (stencil beyond the margin box): appended here, processed by spread_fixup_kernel
    int* exc_count;
    int* exc_list;
    int exc_capacity;
};

// Accumulator tile layout: 16-double rows, row y rotated by 4*y (mod 16).  The 4x4 (x, y) footprint
// of a stencil plane then hits 16 distinct 8-byte banks, and -- unlike an XOR swizzle -- the rotation
// is additive: (x + 4y) & 15 = (R0 + lane constant) & 15 with R0 = x0 + 4*y0 of the stencil origin, so
// a lane needs three integer operations per marker to find its word.
template <int NDIM>
__device__ __forceinline__ int acc_index(int x, int y, int z)
{
    const int xs = (x + 4 * y) & 15;
    if constexpr (NDIM == 3)
        return (z << 8) + (y << 4) + xs;
    else
        return (y << 4) + xs;
}

// brick coordinates from the hierarchical brick id (tile-major, then brick-in-tile)
template <int NDIM>
__device__ __forceinline__ void brick_coords(int brick, const int* nt, int* gb)
{
    if constexpr (NDIM == 3)
    {
        const int tile = brick >> 6;
        const int tx = tile % nt[0], ty = (tile / nt[0]) % nt[1], tz = tile / (nt[0] * nt[1]);
        gb[0] = 4 * tx + (brick & 3);
        gb[1] = 4 * ty + ((brick >> 2) & 3);
        gb[2] = 4 * tz + ((brick >> 4) & 3);
    }
    else
    {
        const int tile = brick >> 4;
        const int tx = tile % nt[0], ty = tile / nt[0];
        gb[0] = 4 * tx + (brick & 3);
        gb[1] = 4 * ty + ((brick >> 2) & 3);
        gb[2] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// stage 1: stencil records, one thread per (marker, dimension, variant)
// ---------------------------------------------------------------------------------------------
template <int NDIM, int K>
__global__ void __launch_bounds__(256) spread_records_kernel(const __grid_constant__ TileParams tp, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    using RL = RecLayout<NDIM, W>;
    constexpr int TASKS = NDIM * 2;
    constexpr int LD = NDIM - 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = args.first + (int)(gid / TASKS);
    const int t = (int)(gid % TASKS);
    if (i >= args.last) return;
    const int d = t >> 1, v = t & 1;
    double* rec = args.records + (long long)i * RL::DOUBLES;
    int* ri = reinterpret_cast<int*>(rec + RL::INTS);
    const unsigned long long key = __ldg(&args.keys[i]);
    int gb[3];
    brick_coords<NDIM>((int)(key >> (args.tie_bits + 2 * NDIM)) - tp.brick_base, tp.nt, gb);
    const int cc = BRICK * gb[d] + (int)((key >> (args.tie_bits + 2 * d)) & 3ull); // binning cell (pp)
    if (v < tp.nvar[d])
    {
        const double xs = __ldg(&args.X[d * args.x_stride + i]);
        const double xr = args.Xraw ? __ldg(&args.Xraw[d * args.x_stride + i]) : xs;
        double w[W];
        int l;
        stencil_1d<K>(xs, xr, tp.xl[d][v], tp.dx[d], l, w);
        const int lo_pp = l + tp.G;
        if (d < LD)
        {
#pragma unroll
            for (int j = 0; j < W; ++j) rec[(d * 2 + v) * W + j] = w[j];
        }
        else
        {
            // last dimension: fold the scaled force of every component that uses this variant
            const long long row = args.src ? (long long)__ldg(&args.src[i]) : (long long)i;
            for (int a = 0; a < args.ncomp; ++a)
            {
                const CompGeom& cg = tp.comp[args.comp0 + a];
                if (cg.var[LD] != v) continue;
                const double f = __ldg(&args.V[cg.vcol * args.v_cstride + row * args.v_istride]) * tp.inv_vol;
#pragma unroll
                for (int j = 0; j < W; ++j) rec[RL::WLF + a * W + j] = w[j] * f;
            }
        }
        ri[d * 2 + v] = lo_pp;
        if ((lo_pp < cc - M || lo_pp + W - 1 > cc + M) && args.exc_list)
        {
            const int slot = atomicAdd(args.exc_count, 1);
            if (slot < args.exc_capacity) args.exc_list[slot] = i;
        }
    }
    else
    {
        ri[d * 2 + v] = -(1 << 20);
    }
    if (v == 0) ri[6 + d] = cc;
    if (t == 0)
    {
        if (NDIM == 2) ri[8] = 0;
        ri[9] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// stage 2: owner-computes tiles
// ---------------------------------------------------------------------------------------------
template <int NDIM, int K, int TZ>
__global__ void __launch_bounds__(SPREAD_THREADS, (NDIM == 3 && TZ <= 8) ? 3 : 2) spread_tile_kernel(const __grid_constant__ TileParams tp, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    using RL = RecLayout<NDIM, W>;
    constexpr int RECD = RL::DOUBLES;
    constexpr int NBR = TILE_BRICKS + (2 * M + BRICK - 1) / BRICK; // bricks per dimension around the tile (x, y)
    constexpr int NBRZ = (NDIM == 3) ? TZ / BRICK + (2 * M + BRICK - 1) / BRICK : 1; // ... and along z (tile is TZ deep)
    constexpr int NC = (BRICK + 2 * M + BRICK - 1) / BRICK;        // colours per dimension
    constexpr int NPTS = (NDIM == 3) ? W * W * W : W * W;
    constexpr int NSLOT = (NPTS + 31) / 32;
    constexpr int TILE_PTS = (NDIM == 3) ? TILE * TILE * TZ : TILE * TILE;
    constexpr int NBRICKS = (NDIM == 3) ? NBR * NBR * NBRZ : NBR * NBR;
    constexpr int NCOL = (NDIM == 3) ? NC * NC * NC : NC * NC;
    constexpr bool FAST4 = (NDIM == 3) && (W == 4); // lane = (ix, iy, half): two adjacent z points per lane
    constexpr int LD = NDIM - 1;                    // the "last" dimension carries the force factor
    constexpr int PF = (SPREAD_BATCH * RECD / 2 + 31) / 32; // 16-byte words per lane per batch

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* acc = reinterpret_cast<double*>(smem_raw);                          // [ncomp][TILE_PTS]
    double* rbuf_all = acc + (size_t)args.ncomp * TILE_PTS;                     // [warp][BATCH][RECD] marker records
    int* rec_all = reinterpret_cast<int*>(rbuf_all + SPREAD_WARPS * SPREAD_BATCH * RECD); // [warp][BATCH][MAXC][2]
    int* brng = rec_all + SPREAD_WARPS * SPREAD_BATCH * SPREAD_MAXC * 2;        // [NBRICKS][2]
    unsigned char* order = reinterpret_cast<unsigned char*>(brng + 2 * NBRICKS); // [NBRICKS] bricks sorted by colour
    __shared__ int any_markers;
    __shared__ int col_start[NCOL + 1];
    __shared__ int claim;                  // next brick (colour-major order) to hand out
    __shared__ unsigned char done[NBRICKS]; // per brick: all its markers have been accumulated

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* rbuf = rbuf_all + warp * (SPREAD_BATCH * RECD);
    int* rec = rec_all + warp * (SPREAD_BATCH * SPREAD_MAXC * 2);

    // which output tile
    int ot[3];
    {
        int r = blockIdx.x;
        ot[0] = tp.ot_lo[0] + r % tp.ot_n[0];
        r /= tp.ot_n[0];
        ot[1] = tp.ot_lo[1] + r % tp.ot_n[1];
        ot[2] = (NDIM == 3) ? tp.ot_lo[2] + r / tp.ot_n[1] : 0;
    }
    int tlo[3]; // pp coordinate of the tile's first point
#pragma unroll
    for (int d = 0; d < 3; ++d) tlo[d] = ((d == 2) ? TZ : TILE) * ot[d] + M;

    // brick ranges of the neighbourhood, colour-sorted brick order, emptiness test
    if (threadIdx.x == 0) any_markers = 0;
    if (threadIdx.x <= NCOL)
    {
        int s = 0; // bricks with colour < threadIdx.x (NBR, NC are compile-time)
        for (int c = 0; c < (int)threadIdx.x; ++c)
        {
            const int c0 = c % NC, c1 = (c / NC) % NC, c2 = (NDIM == 3) ? c / (NC * NC) : 0;
            s += ((NBR - c0 + NC - 1) / NC) * ((NBR - c1 + NC - 1) / NC) * ((NDIM == 3) ? max((NBRZ - c2 + NC - 1) / NC, 0) : 1);
        }
        col_start[threadIdx.x] = s;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NBRICKS; q += SPREAD_THREADS)
    {
        const int lx = q % NBR, ly = (q / NBR) % NBR, lz = (NDIM == 3) ? q / (NBR * NBR) : 0;
        const int bx = TILE_BRICKS * ot[0] + lx, by = TILE_BRICKS * ot[1] + ly, bz = (NDIM == 3) ? (TZ / BRICK) * ot[2] + lz : 0;
        int s = 0, e = 0;
        if (bx < tp.nb[0] && by < tp.nb[1] && bz < tp.nb[2])
        {
            const int id = tp.brick_base + ((NDIM == 3) ? brick_id_3d(bx, by, bz, tp.nt) : brick_id_2d(bx, by, tp.nt));
            s = args.brick_start[id];
            e = args.brick_start[id + 1];
        }
        const int c0 = lx % NC, c1 = ly % NC, c2 = lz % NC;
        const int col = (c2 * NC + c1) * NC + c0;
        const int n0 = (NBR - c0 + NC - 1) / NC, n1 = (NBR - c1 + NC - 1) / NC;
        const int pos = col_start[col] + ((lz / NC) * n1 + ly / NC) * n0 + lx / NC;
        brng[2 * pos] = s; // stored in colour order
        brng[2 * pos + 1] = e;
        order[pos] = (unsigned char)q;
        done[q] = (e > s) ? 0 : 1;
        if (e > s) any_markers = 1;
    }
    __syncthreads();
    if (!any_markers) return;

    for (int q = threadIdx.x; q < args.ncomp * TILE_PTS; q += SPREAD_THREADS) acc[q] = 0.0;

    // Everything this tile will read later -- the records of all its bricks and the f rows of the final
    // `f += tile` -- is requested into L2 now, so that the per-colour stages and the write-out pay an
    // L2 hit instead of a DRAM round trip each (the stages are short and latency-bound).
    for (int q = warp; q < NBRICKS; q += SPREAD_WARPS)
    {
        const int s = brng[2 * q], e = brng[2 * q + 1];
        const char* p0 = reinterpret_cast<const char*>(args.records + (long long)s * RECD);
        const long long bytes = (long long)(e - s) * RECD * 8;
        for (long long o = 128ll * lane; o < bytes; o += 128ll * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
    }
    for (int a = 0; a < args.ncomp; ++a)
    {
        const CompGeom& cg = tp.comp[args.comp0 + a];
        constexpr int ROWS = (NDIM == 3) ? TILE * TZ : TILE;
        for (int row = threadIdx.x; row < ROWS; row += SPREAD_THREADS)
        {
            const int y = row & 15, z = (NDIM == 3) ? row >> 4 : 0;
            const int gi = max(tlo[0] - cg.pp0[0], 0);
            const int gj = tlo[1] + y - cg.pp0[1], gk = (NDIM == 3) ? tlo[2] + z - cg.pp0[2] : 0;
            if (gi < cg.n[0] && gj >= 0 && gj < cg.n[1] && gk >= 0 && gk < cg.n[2])
            {
                const double* p0 = cg.ptr + ((long long)gk * cg.n[1] + gj) * cg.pitch + gi;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p0));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + 15)); // the 16-point row may straddle two lines
            }
        }
    }

    const int ncomp = args.ncomp;
    // A2 role of this lane: (marker, component)
    const int b_m = lane / SPREAD_MAXC, b_a = lane % SPREAD_MAXC;
    const bool b_on = lane < SPREAD_BATCH * SPREAD_MAXC && b_a < ncomp;
    int b_v[3] = { 0, 0, 0 };
    if (b_on)
    {
        const CompGeom& cg = tp.comp[args.comp0 + b_a];
        b_v[0] = cg.var[0];
        b_v[1] = cg.var[1];
        b_v[2] = cg.var[2];
    }
    // phase-B role: weight offsets of each component (which x_lower variant per dimension)
    int wo0[SPREAD_MAXC], wo1[SPREAD_MAXC];
#pragma unroll
    for (int a = 0; a < SPREAD_MAXC; ++a)
    {
        const CompGeom& cg = tp.comp[args.comp0 + (a < ncomp ? a : 0)];
        wo0[a] = (0 * 2 + cg.var[0]) * W;
        wo1[a] = (1 * 2 + cg.var[1]) * W;
    }
    const int l15 = lane & 15, half = lane >> 4;
    int pix[NSLOT], piy[NSLOT], piz[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
    {
        const int q = lane + 32 * s;
        pix[s] = q % W;
        piy[s] = (q / W) % W;
        piz[s] = (NDIM == 3) ? q / (W * W) : 0;
    }
    const double2* recs2 = reinterpret_cast<const double2*>(args.records);
    if (threadIdx.x == 0) claim = 0;
    __syncthreads();

    // Bricks are claimed in colour-major order.  Instead of a CTA barrier between colours, a brick waits
    // only for ITS lower-coloured neighbours (the only bricks whose footprint overlaps its own and that
    // precede it in the fixed summation order): a flag per brick, set when its last marker is done.
    // Claims are handed out in order, so every brick a warp waits for is already owned by a running warp.
    while (true)
    {
        int p = 0;
        if (lane == 0) p = atomicAdd(&claim, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= NBRICKS) break;
        const int bs = brng[2 * p], be = brng[2 * p + 1];
        if (be <= bs) continue; // empty bricks were flagged done at set-up
        const int q = order[p];
        double2 pre[PF];
        {
            const int nv = min(SPREAD_BATCH, be - bs) * (RECD / 2);
            const double2* src = recs2 + (long long)bs * (RECD / 2);
#pragma unroll
            for (int pp = 0; pp < PF; ++pp)
                if (lane + 32 * pp < nv) pre[pp] = __ldg(&src[lane + 32 * pp]);
        }
        {
            const int lx = q % NBR, ly = (q / NBR) % NBR, lz = (NDIM == 3) ? q / (NBR * NBR) : 0;
            const int mycol = ((lz % NC) * NC + ly % NC) * NC + lx % NC;
            // neighbours whose footprint can overlap: index distance <= NC - 1 per dimension
            constexpr int RN = 2 * (NC - 1) + 1;
            constexpr int NNB = (NDIM == 3) ? RN * RN * RN : RN * RN;
            const volatile unsigned char* vd = done;
            for (int n0 = 0; n0 < NNB; n0 += 32)
            {
                const int n = n0 + lane;
                bool need = false;
                int nq = 0;
                if (n < NNB)
                {
                    const int nx = lx + n % RN - (NC - 1), ny = ly + (n / RN) % RN - (NC - 1),
                              nz = (NDIM == 3) ? lz + n / (RN * RN) - (NC - 1) : 0;
                    if (nx >= 0 && nx < NBR && ny >= 0 && ny < NBR && nz >= 0 && nz < NBRZ)
                    {
                        nq = (nz * NBR + ny) * NBR + nx;
                        need = (((nz % NC) * NC + ny % NC) * NC + nx % NC) < mycol;
                    }
                }
                while (!__all_sync(0xffffffffu, !need || vd[nq] != 0))
                {
                }
            }
            __threadfence_block();
        }
        for (int batch = bs; batch < be; batch += SPREAD_BATCH)
        {
                const int nb = min(SPREAD_BATCH, be - batch);
                {
                    const int nv = nb * (RECD / 2);
                    double2* dst = reinterpret_cast<double2*>(rbuf);
#pragma unroll
                    for (int p = 0; p < PF; ++p)
                        if (lane + 32 * p < nv) dst[lane + 32 * p] = pre[p];
                }
                __syncwarp();
                if (batch + SPREAD_BATCH < be)
                {
                    // dense brick: keep one batch in flight
                    const int nv = min(SPREAD_BATCH, be - batch - SPREAD_BATCH) * (RECD / 2);
                    const double2* src = recs2 + (long long)(batch + SPREAD_BATCH) * (RECD / 2);
#pragma unroll
                    for (int p = 0; p < PF; ++p)
                        if (lane + 32 * p < nv) pre[p] = __ldg(&src[lane + 32 * p]);
                }
                // ---- A2: one lane per (marker, component): packed record + last-dim weights * force ----
                if (b_on && b_m < nb)
                {
                    const double* rm = rbuf + b_m * RECD;
                    const int* ri = reinterpret_cast<const int*>(rm + RL::INTS);
                    int o[3] = { 0, 0, 0 };
                    unsigned mk[3] = { 1u, 1u, 1u };
#pragma unroll
                    for (int d = 0; d < NDIM; ++d)
                    {
                        const int lo = ri[d * 2 + b_v[d]], cc = ri[6 + d];
                        const int jlo = max(max(cc - M, tlo[d]) - lo, 0);
                        const int jhi = min(min(cc + M, tlo[d] + ((d == 2) ? TZ : TILE) - 1) - lo, W - 1);
                        mk[d] = (jhi >= jlo) ? (((1u << (jhi + 1)) - 1u) & ~((1u << jlo) - 1u)) : 0u;
                        o[d] = lo - tlo[d];
                    }
                    unsigned w0r;
                    if constexpr (FAST4)
                    {
                        unsigned mxy = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) mxy |= ((mk[1] >> j) & 1u) ? (mk[0] << (4 * j)) : 0u;
                        w0r = mxy | (mk[2] << 16);
                    }
                    else
                    {
                        w0r = mk[0] | (mk[1] << 8) | (mk[2] << 16);
                    }
                    const bool empty = (mk[0] == 0) || (mk[1] == 0) || (mk[2] == 0);
                    w0r = empty ? 0u : (w0r | ((unsigned)((o[0] + 4 * o[1]) & 15) << 24) | 0x80000000u);
                    rec[(b_m * SPREAD_MAXC + b_a) * 2 + 0] = (int)w0r;
                    rec[(b_m * SPREAD_MAXC + b_a) * 2 + 1] = (NDIM == 3) ? (o[2] * 256 + o[1] * 16) : (o[1] * 16);
                }
                __syncwarp();
                // ---- phase B: markers one after another, lanes over the stencil points ----
                for (int m = 0; m < nb; ++m)
                {
                    const double* wm = rbuf + m * RECD;
#pragma unroll
                    for (int a = 0; a < SPREAD_MAXC; ++a)
                    {
                        if (a >= ncomp) break;
                        const int2 r = *reinterpret_cast<const int2*>(&rec[(m * SPREAD_MAXC + a) * 2]);
                        if (r.x >= 0) continue; // nothing of this marker lands in the tile (warp-uniform)
                        double* acc_a = acc + a * TILE_PTS;
                        const int rot = (r.x >> 24) & 15;
                        if constexpr (FAST4)
                        {
                            // lane = (ix, iy) = l15, z points 2*half and 2*half + 1
                            const int idx = r.y + (half << 9) + ((l15 >> 2) << 4) + ((rot + l15) & 15);
                            const double wxy = wm[wo0[a] + (l15 & 3)] * wm[wo1[a] + (l15 >> 2)];
                            const double2 wz = *reinterpret_cast<const double2*>(&wm[RL::WLF + a * W + 2 * half]);
                            const bool okxy = (r.x >> l15) & 1;
                            const bool ok0 = okxy && ((r.x >> (16 + 2 * half)) & 1);
                            const bool ok1 = okxy && ((r.x >> (17 + 2 * half)) & 1);
                            if (ok0) acc_a[idx] += wxy * wz.x;
                            if (ok1) acc_a[idx + 256] += wxy * wz.y;
                        }
                        else
                        {
#pragma unroll
                            for (int s = 0; s < NSLOT; ++s)
                            {
                                const bool active = (NPTS % 32 == 0) || (lane + 32 * s < NPTS);
                                if (!active) continue;
                                bool ok = ((r.x >> pix[s]) & (r.x >> (8 + piy[s])) & 1) != 0;
                                double wv = wm[wo0[a] + pix[s]];
                                int idx = r.y + (piy[s] << 4) + ((rot + pix[s] + 4 * piy[s]) & 15);
                                if constexpr (NDIM == 3)
                                {
                                    ok = ok && ((r.x >> (16 + piz[s])) & 1);
                                    wv *= wm[wo1[a] + piy[s]] * wm[RL::WLF + a * W + piz[s]];
                                    idx += piz[s] << 8;
                                }
                                else
                                {
                                    wv *= wm[RL::WLF + a * W + piy[s]];
                                }
                                if (ok) acc_a[idx] += wv;
                            }
                        }
                    }
                    __syncwarp();
                }
        }
        __syncwarp();
        if (lane == 0)
        {
            __threadfence_block();
            done[q] = 1;
        }
    }
    __syncthreads();


    // ---- write-out: f += tile (coalesced along x), dropping points outside the array ----
    // thread = (x, y); it walks the z column (3D) with a constant pointer / index stride
    {
        const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
        constexpr int NZ = (NDIM == 3) ? TZ : 1;
        const int sbase = (y << 4) + ((x + 4 * y) & 15);
        for (int a = 0; a < ncomp; ++a)
        {
            const CompGeom& cg = tp.comp[args.comp0 + a];
            const double* acc_a = acc + a * TILE_PTS + sbase;
            const int gi = tlo[0] + x - cg.pp0[0], gj = tlo[1] + y - cg.pp0[1];
            const int gk0 = (NDIM == 3) ? tlo[2] - cg.pp0[2] : 0;
            const bool okxy = gi >= 0 && gi < cg.n[0] && gj >= 0 && gj < cg.n[1];
            const long long zstride = (long long)cg.n[1] * cg.pitch;
            double* p0 = cg.ptr + ((long long)gk0 * cg.n[1] + gj) * cg.pitch + gi;
            const int zlo = max(0, -gk0), zhi = min(NZ, cg.n[2] - gk0); // valid z range of this tile
            if (!okxy) continue;
#pragma unroll
            for (int z0 = 0; z0 < NZ; z0 += 8)
            {
                double v[8], old[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                {
                    const int z = z0 + r;
                    v[r] = (z < NZ && z >= zlo && z < zhi) ? acc_a[z << 8] : 0.0;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) old[r] = (v[r] != 0.0) ? p0[(z0 + r) * zstride] : 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (v[r] != 0.0) p0[(z0 + r) * zstride] = old[r] + v[r];
            }
        }
    }
}

// Fix-up for the (practically never occurring) markers whose stencil reaches beyond M cells from
// their binning cell: one thread, sorted-position order, only the points OUTSIDE the margin box
// (the tile kernel did the ones inside).  Deterministic by construction.
template <int NDIM, int K>
__global__ void spread_fixup_kernel(const __grid_constant__ TileParams tp, SpreadArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = *args.exc_count;
    if (n <= 0) return;
    if (n > args.exc_capacity) n = args.exc_capacity;
    // insertion sort of the exception list by sorted position (tiny)
    for (int a = 1; a < n; ++a)
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
        const int i = args.exc_list[e];
        if (e > 0 && args.exc_list[e - 1] == i) continue;
        const uint64_t key = args.keys[i];
        const long long brick_global = (long long)(key >> (args.tie_bits + 2 * NDIM));
        const int brick = (int)(brick_global - tp.brick_base);
        // decode hierarchical brick id -> brick coordinates
        int gb[3] = { 0, 0, 0 };
        if (NDIM == 3)
        {
            const int tile = brick >> 6;
            const int tx = tile % tp.nt[0], ty = (tile / tp.nt[0]) % tp.nt[1], tz = tile / (tp.nt[0] * tp.nt[1]);
            gb[0] = 4 * tx + (brick & 3);
            gb[1] = 4 * ty + ((brick >> 2) & 3);
            gb[2] = 4 * tz + ((brick >> 4) & 3);
        }
        else
        {
            const int tile = brick >> 4;
            const int tx = tile % tp.nt[0], ty = tile / tp.nt[0];
            gb[0] = 4 * tx + (brick & 3);
            gb[1] = 4 * ty + ((brick >> 2) & 3);
        }
        const long long row = args.src ? (long long)args.src[i] : (long long)i;
        for (int a = 0; a < args.ncomp; ++a)
        {
            const CompGeom& cg = tp.comp[args.comp0 + a];
            double w[3][W];
            int lo[3] = { 0, 0, 0 }, cc[3] = { 0, 0, 0 };
            for (int d = 0; d < NDIM; ++d)
            {
                const double xs = args.X[d * args.x_stride + i];
                const double xr = args.Xraw ? args.Xraw[d * args.x_stride + i] : xs;
                int l;
                stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w[d]);
                lo[d] = l + tp.G;
                cc[d] = BRICK * gb[d] + (int)((key >> (args.tie_bits + 2 * d)) & 3ull);
            }
            const double f = args.V[cg.vcol * args.v_cstride + row * args.v_istride] * tp.inv_vol;
            const int KW = (NDIM == 3) ? W : 1;
            for (int k = 0; k < KW; ++k)
                for (int j = 0; j < W; ++j)
                    for (int ii = 0; ii < W; ++ii)
                    {
                        const int px = lo[0] + ii, py = lo[1] + j, pz = (NDIM == 3) ? lo[2] + k : 0;
                        bool inside = (px >= cc[0] - M && px <= cc[0] + M) && (py >= cc[1] - M && py <= cc[1] + M);
                        if (NDIM == 3) inside = inside && (pz >= cc[2] - M && pz <= cc[2] + M);
                        if (inside) continue; // done by the tile kernel
                        const int gi = px - cg.pp0[0], gj = py - cg.pp0[1], gk = (NDIM == 3) ? pz - cg.pp0[2] : 0;
                        if (gi < 0 || gi >= cg.n[0] || gj < 0 || gj >= cg.n[1] || gk < 0 || gk >= cg.n[2]) continue;
                        double wv = w[0][ii] * w[1][j];
                        if (NDIM == 3) wv *= w[2][k];
                        cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi] += wv * f;
                    }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int* g_exc_buf = nullptr; // [1 + capacity] per process (device); tiny
constexpr int EXC_CAPACITY = 4096;
static double* g_rec_buf = nullptr; // marker records, grown on demand
static size_t g_rec_cap = 0;

template <int NDIM, int K, int TZ>
static cudaError_t launch_spread_tz(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    using RL = RecLayout<NDIM, W>;
    constexpr int NBR = TILE_BRICKS + (2 * M + BRICK - 1) / BRICK;
    constexpr int NBRZ = (NDIM == 3) ? TZ / BRICK + (2 * M + BRICK - 1) / BRICK : 1;
    constexpr int TILE_PTS = (NDIM == 3) ? TILE * TILE * TZ : TILE * TILE;
    constexpr int NBRICKS = (NDIM == 3) ? NBR * NBR * NBRZ : NBR * NBR;
    cudaError_t e;
    if (!g_exc_buf)
    {
        if ((e = cudaMalloc(&g_exc_buf, sizeof(int) * (1 + EXC_CAPACITY))) != cudaSuccess) return e;
    }
    const size_t need = (size_t)std::max(bins.n_entries, 1) * RL::DOUBLES * sizeof(double);
    if (need > g_rec_cap)
    {
        if (g_rec_buf) cudaFree(g_rec_buf);
        g_rec_buf = nullptr;
        g_rec_cap = 0;
        if ((e = cudaMalloc(&g_rec_buf, need + need / 8)) != cudaSuccess)
        {
            err = "cudaMalloc(marker records) failed";
            return e;
        }
        g_rec_cap = need + need / 8;
    }
    if ((e = cudaMemsetAsync(g_exc_buf, 0, sizeof(int), L.stream)) != cudaSuccess) return e;
    SpreadArgs args;
    args.brick_start = bins.brick_start;
    args.keys = bins.keys[bins.sorted_in];
    args.tie_bits = bins.tie_bits;
    args.X = mv.X;
    args.Xraw = mv.Xraw;
    args.x_stride = mv.x_stride;
    args.V = mv.V;
    args.v_cstride = mv.v_cstride;
    args.v_istride = mv.v_istride;
    args.src = mv.src;
    args.records = g_rec_buf;
    args.exc_count = g_exc_buf;
    args.exc_list = g_exc_buf + 1;
    args.exc_capacity = EXC_CAPACITY;
    // sorted positions of this patch's markers (read back when the bins were built)
    args.first = args.last = 0;
    for (size_t p = 0; p < bins.range_base.size(); ++p)
        if (bins.range_base[p] == tp.brick_base)
        {
            args.first = bins.range_first[p];
            args.last = bins.range_last[p];
        }
    if (args.last <= args.first) return cudaSuccess;
    // output tiles: tile a covers the points pp in [16a + M, 16a + M + 16); cover every array point
    TileParams tpl = tp;
    for (int d = 0; d < 3; ++d)
    {
        tpl.ot_lo[d] = 0;
        tpl.ot_n[d] = 1;
        if (d >= NDIM) continue;
        int ppmin = 1 << 30, ppmax = -(1 << 30);
        for (int a = 0; a < tp.ncomp; ++a)
        {
            ppmin = std::min(ppmin, tp.comp[a].pp0[d]);
            ppmax = std::max(ppmax, tp.comp[a].pp0[d] + tp.comp[a].n[d] - 1);
        }
        auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
        const int td = (d == 2) ? TZ : TILE;
        const int alo = std::max(0, fdiv(ppmin - M, td));
        const int ahi = fdiv(ppmax - M, td);
        tpl.ot_lo[d] = alo;
        tpl.ot_n[d] = std::max(0, ahi - alo + 1);
    }
    const int ntiles = tpl.ot_n[0] * tpl.ot_n[1] * tpl.ot_n[2];
    if (ntiles <= 0) return cudaSuccess;
    auto rfn = spread_records_kernel<NDIM, K>;
    auto kfn = spread_tile_kernel<NDIM, K, TZ>;
    auto ffn = spread_fixup_kernel<NDIM, K>;
    for (int c0 = 0; c0 < tp.ncomp; c0 += SPREAD_MAXC)
    {
        args.comp0 = c0;
        args.ncomp = (tp.ncomp - c0 < SPREAD_MAXC) ? tp.ncomp - c0 : SPREAD_MAXC;
        const size_t smem = sizeof(double) * ((size_t)args.ncomp * TILE_PTS + SPREAD_WARPS * SPREAD_BATCH * RL::DOUBLES) +
                            sizeof(int) * (SPREAD_WARPS * SPREAD_BATCH * SPREAD_MAXC * 2 + 2 * NBRICKS) + ((NBRICKS + 15) / 16) * 16;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
        {
            err = "cudaFuncSetAttribute(spread) failed";
            return e;
        }
        const long long ntask = (long long)(args.last - args.first) * NDIM * 2;
        rfn<<<(unsigned)((ntask + 255) / 256), 256, 0, L.stream>>>(tpl, args);
        kfn<<<ntiles, SPREAD_THREADS, smem, L.stream>>>(tpl, args);
        ffn<<<1, 32, 0, L.stream>>>(tpl, args);
        L.launches += 3;
        if (c0 + SPREAD_MAXC < tp.ncomp)
            if ((e = cudaMemsetAsync(g_exc_buf, 0, sizeof(int), L.stream)) != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// z depth of the output tile: 8 keeps three CTAs (24 warps) resident per SM for the 4-point kernels; the
// kernel is latency-bound, so the extra resident warps outweigh the larger marker neighbourhood
template <int NDIM, int K>
static cudaError_t launch_spread_t(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    static const char* env = getenv("IBK_SPREAD_TZ");
    const int tz = env ? atoi(env) : ((NDIM == 3 && KTraits<K>::W <= 4) ? 8 : 16);
    if (NDIM == 3 && tz == 8) return launch_spread_tz<NDIM, K, 8>(L, tp, bins, mv, err);
    return launch_spread_tz<NDIM, K, 16>(L, tp, bins, mv, err);
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
