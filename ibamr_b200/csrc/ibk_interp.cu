// ibk_interp.cu -- velocity interpolation (grid -> markers) for sm_100a.
//
// Replaces lagrangian_<kernel>_interp{2,3}d (ibtk/src/lagrangian/fortran/
// lagrangian_interaction3d.f.m4:1203-1334 ib_4, :2178-2375 ib_6, :2591-2693 bspline_3,
// :2814-2923 bspline_4, :493-607 piecewise_linear; 2D twins in lagrangian_interaction2d.f.m4) and
// the per-axis loop LEInteractor wraps around them (LEInteractor.cpp:2454-2486).
//
// One CTA per marker tile (16^ndim cells; its markers are one contiguous range of the sorted
// storage).  For every component the CTA stages the tile's grid neighbourhood
// (16 + 2M)^ndim points, M = kernel reach, into shared memory with ONE TMA box copy
// (cp.async.bulk.tensor, out-of-bounds elements are zero-filled, which is exactly the reference's
// "clip the stencil to the ghost box"), waits on an mbarrier, and then each thread gathers the
// tensor-product stencil of one marker from shared memory.  A marker whose stencil is not inside
// the staged box (only possible when its binning cell and its stencil origin disagree by a
// rounding) takes a clipped global-memory path, so results never depend on the staging.
// Arrays that do not satisfy TMA's 16-byte pitch/base alignment (the raw B4 seam on dense
// reference-layout arrays) are staged with ordinary coalesced loads instead.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ibk_engine.h"
#include "ibk_tma.h"

namespace ibk
{
struct InterpArgs
{
    const int* brick_start;
    const double* X;
    const double* Xraw;
    long long x_stride;
    double* V;
    long long v_cstride, v_istride;
    const uint32_t* src;
    unsigned tma_mask; // bit a set: component a is staged by TMA
    int part, sel_lo[3], sel_hi[3]; // MarkerView's tile selection
};

constexpr int INTERP_THREADS = 256;
// A tile of the uniform benchmark holds 256 +- 16 markers: interp_rot_kernel takes 320 threads (20 half-warps x 8 markers
// per parity list) so that one pass covers a tile.  (Measured: 320 threads alone do nothing for the plain tile kernel, 1.56 ms;
// 288 threads for this one: 1.31 ms against 1.26 ms; one CTA per (tile, component): 1.38 ms.)
constexpr int INTERP_THREADS_WIDE = 320;
constexpr bool INTERP_ROT_DEFAULT = true; // measured: 1.26 ms against 1.54 ms on the C5 shard (IBK_INTERP_ROT=0: the plain tile kernel)

// (min CTAs per SM = what the staged box allows: without it ptxas takes 110 registers and only two CTAs fit)
template <int NDIM, int K, int NT>
__global__ void __launch_bounds__(NT, (KTraits<K>::M <= 2) ? 3 : (KTraits<K>::M == 3) ? 2 : 1)
    interp_tile_kernel(const __grid_constant__ TileParams tp, const __grid_constant__ TmaMapSet maps, InterpArgs args)
{
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int S = TILE + 2 * M;
    // TMA needs the box to start on a 16-byte boundary: with 8-byte elements the innermost start
    // coordinate must be even.  The box is therefore SX = S + 2 wide in x and starts at the even
    // coordinate at or below the first needed point (measured on B200: an odd start traps).
    constexpr int SX = S + 2;
    constexpr int BRICKS_PER_TILE = (NDIM == 3) ? 64 : 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* su = reinterpret_cast<double*>(smem_raw);
    __shared__ uint64_t bar;

    const int tile = blockIdx.x;
    int t[3];
    {
        int r = tile;
        t[0] = r % tp.nt[0];
        r /= tp.nt[0];
        t[1] = r % tp.nt[1];
        t[2] = r / tp.nt[1];
    }
    if (args.part)
    {
        bool in = true;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) in = in && t[d] >= args.sel_lo[d] && t[d] <= args.sel_hi[d];
        if ((args.part == 1) != in) return;
    }
    const int b0 = tp.brick_base + tile * BRICKS_PER_TILE;
    const int s0 = args.brick_start[b0];
    const int s1 = args.brick_start[b0 + BRICKS_PER_TILE];
    if (s0 >= s1) return;

    // pp coordinate of the first staged point per dimension
    int sp0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) sp0[d] = TILE * t[d] - M;

    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    // The boxes of the later components are requested into L2 now (TMA prefetch), so that their loads
    // pay an L2 hit instead of a DRAM round trip while the CTA sits at the mbarrier.
    if (threadIdx.x == 0)
    {
        for (int a = 1; a < tp.ncomp; ++a)
        {
            if (!((args.tma_mask >> a) & 1u)) continue;
            const CompGeom& cg = tp.comp[a];
            int c0 = sp0[0] - cg.pp0[0];
            c0 -= (c0 & 1);
            const int c1 = sp0[1] - cg.pp0[1], c2 = sp0[2] - cg.pp0[2];
            if constexpr (NDIM == 3)
                asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&maps.m[a]), "r"(c0), "r"(c1), "r"(c2) : "memory");
            else
                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&maps.m[a]), "r"(c0), "r"(c1) : "memory");
        }
    }
    // positions of this thread's first marker stay in registers across the component loop
    const int i_first = s0 + threadIdx.x;
    double xs0[NDIM], xr0[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
    {
        xs0[d] = (i_first < s1) ? args.X[d * args.x_stride + i_first] : 0.0;
        xr0[d] = (i_first < s1 && args.Xraw) ? args.Xraw[d * args.x_stride + i_first] : xs0[d];
    }

    for (int a = 0; a < tp.ncomp; ++a)
    {
        const CompGeom& cg = tp.comp[a];
        // element coordinates of the staged box's first point
        int e0[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) e0[d] = sp0[d] - cg.pp0[d];
        const int xodd = e0[0] & 1; // two's complement: also right for negative coordinates
        e0[0] -= xodd;
        const int sx0 = sp0[0] - xodd; // pp coordinate of the first staged column
        const bool use_tma = (args.tma_mask >> a) & 1u;
        if (use_tma)
        {
            if (threadIdx.x == 0)
            {
                constexpr uint32_t bytes = (NDIM == 3 ? SX * S * S : SX * S) * sizeof(double);
                fence_proxy_async_smem();
                mbar_expect_tx(&bar, bytes);
                if constexpr (NDIM == 3)
                    tma_load_3d(su, &maps.m[a], &bar, e0[0], e0[1], e0[2]);
                else
                    tma_load_2d(su, &maps.m[a], &bar, e0[0], e0[1]);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        else
        {
            constexpr int NPTS = (NDIM == 3) ? SX * S * S : SX * S;
            for (int q = threadIdx.x; q < NPTS; q += NT)
            {
                const int i = q % SX;
                const int j = (q / SX) % S;
                const int k = (NDIM == 3) ? q / (SX * S) : 0;
                const int gi = e0[0] + i, gj = e0[1] + j, gk = (NDIM == 3) ? e0[2] + k : 0;
                double v = 0.0;
                if (gi >= 0 && gi < cg.n[0] && gj >= 0 && gj < cg.n[1] && gk >= 0 && gk < cg.n[2])
                    v = cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi];
                su[q] = v;
            }
            __syncthreads();
        }

        for (int i = s0 + threadIdx.x; i < s1; i += NT)
        {
            double w[NDIM][W];
            int lo[NDIM]; // stencil origin in pp coordinates
            bool staged = true;
#pragma unroll
            for (int d = 0; d < NDIM; ++d)
            {
                const double xs = (i == i_first) ? xs0[d] : args.X[d * args.x_stride + i];
                const double xr = (i == i_first) ? xr0[d] : (args.Xraw ? args.Xraw[d * args.x_stride + i] : xs);
                int l;
                stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w[d], d == cg.axis);
                lo[d] = l + tp.G;
                staged = staged && (lo[d] >= sp0[d]) && (lo[d] + W <= sp0[d] + S);
            }
            double acc = 0.0;
            if (staged)
            {
                if constexpr (NDIM == 3)
                {
                    const double* base = su + ((lo[2] - sp0[2]) * S + (lo[1] - sp0[1])) * SX + (lo[0] - sx0);
#pragma unroll
                    for (int k = 0; k < W; ++k)
#pragma unroll
                        for (int j = 0; j < W; ++j)
                        {
                            // the row's x sum first, then its (y, z) weight (the terms of w0 (w1 w2) u, associated
                            // differently: W + 1 instead of 2 W + 1 fp64 instructions per row)
                            const double wyz = w[1][j] * w[2][k];
                            double rs = w[0][0] * base[(k * S + j) * SX];
#pragma unroll
                            for (int ii = 1; ii < W; ++ii) rs += w[0][ii] * base[(k * S + j) * SX + ii];
                            acc += wyz * rs;
                        }
                }
                else
                {
                    const double* base = su + (lo[1] - sp0[1]) * SX + (lo[0] - sx0);
#pragma unroll
                    for (int j = 0; j < W; ++j)
#pragma unroll
                        for (int ii = 0; ii < W; ++ii) acc += (w[0][ii] * w[1][j]) * base[j * SX + ii];
                }
            }
            else
            {
                // clipped global path (3d.f.m4:1309-1327)
                constexpr int KW = (NDIM == 3) ? W : 1;
#pragma unroll
                for (int k = 0; k < KW; ++k)
                {
                    const int gk = (NDIM == 3) ? lo[NDIM - 1] + k - cg.pp0[2] : 0;
                    if (gk < 0 || gk >= cg.n[2]) continue;
#pragma unroll
                    for (int j = 0; j < W; ++j)
                    {
                        const int gj = lo[1] + j - cg.pp0[1];
                        if (gj < 0 || gj >= cg.n[1]) continue;
                        const double wyz = (NDIM == 3) ? w[1][j] * w[NDIM - 1][k] : w[1][j];
#pragma unroll
                        for (int ii = 0; ii < W; ++ii)
                        {
                            const int gi = lo[0] + ii - cg.pp0[0];
                            if (gi < 0 || gi >= cg.n[0]) continue;
                            acc += (w[0][ii] * wyz) * cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi];
                        }
                    }
                }
            }
            const long long row = args.src ? (long long)args.src[i] : (long long)i;
            args.V[cg.vcol * args.v_cstride + row * args.v_istride] = acc;
        }
        __syncthreads(); // everyone is done with su before the next component overwrites it
    }
}


// ---------------------------------------------------------------------------------------------
// Bank-conflict-free gather (3D, 4-point kernels with M = 2: staged box 22 x 20 x 20).
// A thread gathers the 64 stencil values of one marker with the same 64 offsets as every other lane, so the bank
// pattern of all 64 loads is decided by the lanes' box origins: random origins cost 2.9 wavefronts per half-warp load
// (ncu: 300 M shared wavefronts against 104 M ideal).  Two facts remove them:
//  * rows (j, k) of the box add 6 j + 8 k (mod 16) to the 8-byte bank: row q = 4 k + j adds 6 q.  A lane that visits
//    its 16 rows in the ROTATED order q + r therefore shifts all its loads by 6 r, i.e. to any bank of the same parity
//    (6 r runs over the even residues for r = 0..7), and keeps that shift for all 64 loads.
//  * the parity of a lane's bank is the parity of its x origin.  The markers of a chunk are split into an even and an
//    odd list (deterministic ballot scan) and a half-warp takes 8 of each: lane l aims at bank 2 (l & 7) + (l >> 3),
//    so the 16 lanes of a half-warp always hit 16 different banks: one wavefront per half-warp load.
// The rotation only reorders the 16 row sums of a marker; which rotation a marker gets depends on its rank in its
// parity list, i.e. on the tile's markers alone: results are reproducible run to run.
// ---------------------------------------------------------------------------------------------
template <int K, int NT>
__global__ void __launch_bounds__(NT, 3)
    interp_rot_kernel(const __grid_constant__ TileParams tp, const __grid_constant__ TmaMapSet maps, InterpArgs args)
{
    constexpr int NDIM = 3;
    constexpr int W = KTraits<K>::W;
    constexpr int M = KTraits<K>::M;
    constexpr int S = TILE + 2 * M;
    constexpr int SX = S + 2;
    static_assert(W == 4 && (SX % 16) == 6 && ((S * SX) % 16) == 8, "row q = 4 k + j shifts the bank by 6 q");
    constexpr int NWARP = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* su = reinterpret_cast<double*>(smem_raw);
    __shared__ uint64_t bar;
    __shared__ unsigned short plist[2][NT]; // chunk-local marker numbers with an even / odd x origin
    __shared__ int wcnt[2][NWARP];
    __shared__ int ntot[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x;
    int t[3];
    {
        int r = tile;
        t[0] = r % tp.nt[0];
        r /= tp.nt[0];
        t[1] = r % tp.nt[1];
        t[2] = r / tp.nt[1];
    }
    if (args.part)
    {
        bool in = true;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) in = in && t[d] >= args.sel_lo[d] && t[d] <= args.sel_hi[d];
        if ((args.part == 1) != in) return;
    }
    const int b0 = tp.brick_base + tile * 64;
    const int s0 = args.brick_start[b0];
    const int s1 = args.brick_start[b0 + 64];
    if (s0 >= s1) return;
    int sp0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) sp0[d] = TILE * t[d] - M;
    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    if (threadIdx.x == 0)
    {
        for (int a = 1; a < tp.ncomp; ++a)
        {
            const CompGeom& cg = tp.comp[a];
            int c0 = sp0[0] - cg.pp0[0];
            c0 -= (c0 & 1);
            const int c1 = sp0[1] - cg.pp0[1], c2 = sp0[2] - cg.pp0[2];
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&maps.m[a]), "r"(c0), "r"(c1), "r"(c2) : "memory");
        }
    }
    for (int a = 0; a < tp.ncomp; ++a)
    {
        const CompGeom& cg = tp.comp[a];
        int e0[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) e0[d] = sp0[d] - cg.pp0[d];
        const int xodd = e0[0] & 1;
        e0[0] -= xodd;
        const int sx0 = sp0[0] - xodd; // pp coordinate of the first staged column
        if (threadIdx.x == 0)
        {
            constexpr uint32_t bytes = SX * S * S * sizeof(double);
            fence_proxy_async_smem();
            mbar_expect_tx(&bar, bytes);
            tma_load_3d(su, &maps.m[a], &bar, e0[0], e0[1], e0[2]);
        }
        bool waited = false;
        // components with the same x geometry (all but the x-normal one) have the same parity lists
        const bool same_lists = a > 0 && s1 - s0 <= NT && cg.var[0] == tp.comp[a - 1].var[0] && cg.pp0[0] == tp.comp[a - 1].pp0[0] &&
                                (0 == cg.axis) == (0 == tp.comp[a - 1].axis);
        for (int chunk = s0; chunk < s1; chunk += NT)
        {
            // ---- the chunk's markers by the parity of their x origin, in storage order (ballot scan: deterministic)
            const int i1 = chunk + threadIdx.x;
            const bool valid = i1 < s1;
            int par = 0;
            if (!same_lists)
            {
                if (valid)
                {
                    if (a == 0) // the gather reads y and z through the lists (uncoalesced): have them in L1 by then
                    {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(args.X + args.x_stride + i1));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(args.X + 2 * args.x_stride + i1));
                    }
                    const double xs = args.X[i1];
                    const double xr = args.Xraw ? args.Xraw[i1] : xs;
                    double wtmp[W];
                    int l;
                    stencil_1d<K>(xs, xr, tp.xl[0][cg.var[0]], tp.dx[0], l, wtmp, 0 == cg.axis);
                    par = (l + tp.G - sx0) & 1;
                }
                const unsigned mE = __ballot_sync(0xffffffffu, valid && par == 0), mO = __ballot_sync(0xffffffffu, valid && par == 1);
                if (lane == 0)
                {
                    wcnt[0][warp] = __popc(mE);
                    wcnt[1][warp] = __popc(mO);
                }
                __syncthreads();
                int nE = 0, nO = 0, bE = 0, bO = 0;
    #pragma unroll
                for (int w = 0; w < NWARP; ++w)
                {
                    if (w == warp)
                    {
                        bE = nE;
                        bO = nO;
                    }
                    nE += wcnt[0][w];
                    nO += wcnt[1][w];
                }
                // (these totals go to ntot[] below; the gather reads them from there)
                if (valid)
                {
                    const unsigned lt = (1u << lane) - 1u;
                    if (par == 0)
                        plist[0][bE + __popc(mE & lt)] = (unsigned short)threadIdx.x;
                    else
                        plist[1][bO + __popc(mO & lt)] = (unsigned short)threadIdx.x;
                }
                if (threadIdx.x == 0)
                {
                    ntot[0] = nE;
                    ntot[1] = nO;
                }
                __syncthreads();
            }
            const int nE = ntot[0], nO = ntot[1];
            if (!waited)
            {
                mbar_wait(&bar, phase);
                phase ^= 1;
                waited = true;
            }
            // ---- gather: a half-warp takes 8 even and 8 odd markers
            const int hl = threadIdx.x & 15, grp = hl >> 3;
            const int ngrp = grp ? nO : nE;
            const int nmax = max(nE, nO);
            for (int slot = threadIdx.x >> 4; slot * 8 < nmax; slot += NT / 16)
            {
                const int idx = slot * 8 + (hl & 7);
                if (idx >= ngrp) continue;
                const int i = chunk + plist[grp][idx];
                double w[NDIM][W];
                int lo[NDIM];
                bool staged = true;
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                {
                    const double xs = args.X[d * args.x_stride + i];
                    const double xr = args.Xraw ? args.Xraw[d * args.x_stride + i] : xs;
                    int l;
                    stencil_1d<K>(xs, xr, tp.xl[d][cg.var[d]], tp.dx[d], l, w[d], d == cg.axis);
                    lo[d] = l + tp.G;
                    staged = staged && (lo[d] >= sp0[d]) && (lo[d] + W <= sp0[d] + S);
                }
                double acc = 0.0;
                if (staged)
                {
                    const int b = ((lo[2] - sp0[2]) * S + (lo[1] - sp0[1])) * SX + (lo[0] - sx0);
                    const int tgt = 2 * (hl & 7) + grp; // this lane's bank; b has the parity of grp
                    const int r = (3 * (((tgt - b) & 15) >> 1)) & 7;
                    const int rj = r & 3, rk = r >> 2;
                    auto sel4 = [](const double* v, int q) { return q == 0 ? v[0] : q == 1 ? v[1] : q == 2 ? v[2] : v[3]; };
                    double w1r[W], w2r[W];
                    int jo[W], ko[W];
                    bool cy[W];
#pragma unroll
                    for (int j = 0; j < W; ++j)
                    {
                        const int jj = (j + rj) & 3, kk = (j + rk) & 3;
                        w1r[j] = sel4(w[1], jj);
                        jo[j] = jj * SX;
                        cy[j] = (j + rj) >= 4;
                        w2r[j] = sel4(w[2], kk);
                        ko[j] = kk * (S * SX);
                    }
                    const double* bp = su + b;
                    double pacc[W] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
                    for (int k = 0; k < W; ++k)
#pragma unroll
                        for (int j = 0; j < W; ++j)
                        {
                            // static step 4 k + j visits row 4 k + j + r: (j + rj) & 3 in plane (k + rk + carry) & 3
                            const double w2s = cy[j] ? w2r[(k + 1) & 3] : w2r[k];
                            const int kos = cy[j] ? ko[(k + 1) & 3] : ko[k];
                            const double wyz = w1r[j] * w2s;
                            const double* row = bp + jo[j] + kos;
                            // the row's x sum first, then its (y, z) weight: 5 instead of 9 fp64 instructions per row
                            // (same terms as w0 (w1 w2) u, associated differently: rounding-level difference only)
                            double rs = w[0][0] * row[0];
#pragma unroll
                            for (int ii = 1; ii < W; ++ii) rs += w[0][ii] * row[ii];
                            pacc[j] += wyz * rs;
                        }
                    acc = (pacc[0] + pacc[1]) + (pacc[2] + pacc[3]); // four chains of 16 instead of one of 64
                }
                else
                {
                    // clipped global path (3d.f.m4:1309-1327)
#pragma unroll
                    for (int k = 0; k < W; ++k)
                    {
                        const int gk = lo[2] + k - cg.pp0[2];
                        if (gk < 0 || gk >= cg.n[2]) continue;
#pragma unroll
                        for (int j = 0; j < W; ++j)
                        {
                            const int gj = lo[1] + j - cg.pp0[1];
                            if (gj < 0 || gj >= cg.n[1]) continue;
                            const double wyz = w[1][j] * w[2][k];
#pragma unroll
                            for (int ii = 0; ii < W; ++ii)
                            {
                                const int gi = lo[0] + ii - cg.pp0[0];
                                if (gi < 0 || gi >= cg.n[0]) continue;
                                acc += (w[0][ii] * wyz) * cg.ptr[((long long)gk * cg.n[1] + gj) * cg.pitch + gi];
                            }
                        }
                    }
                }
                const long long row = args.src ? (long long)args.src[i] : (long long)i;
                args.V[cg.vcol * args.v_cstride + row * args.v_istride] = acc;
            }
            __syncthreads(); // the lists and (after the last chunk) su are free again
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

bool make_tensor_map(CUtensorMap* m, const CompGeom& cg, int ndim, unsigned bx, unsigned by, unsigned bz, int promo)
{
    if (((uintptr_t)cg.ptr % 16 != 0) || ((cg.pitch * 8) % 16 != 0)) return false;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = { (cuuint64_t)cg.n[0], (cuuint64_t)cg.n[1], (cuuint64_t)cg.n[2] };
    cuuint64_t strides[2] = { (cuuint64_t)cg.pitch * 8ull, (cuuint64_t)cg.pitch * 8ull * (cuuint64_t)cg.n[1] };
    cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)ndim, (void*)cg.ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                     promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int NDIM, int K>
static cudaError_t launch_interp_t(Launcher& L, const TileParams& tp, const Bins& bins, const MarkerView& mv, std::string& err)
{
    constexpr int M = KTraits<K>::M;
    constexpr int S = TILE + 2 * M;
    const size_t smem = sizeof(double) * (size_t)(NDIM == 3 ? (S + 2) * S * S : (S + 2) * S);
    TmaMapSet maps;
    std::memset(&maps, 0, sizeof(maps));
    InterpArgs args;
    args.brick_start = bins.brick_start;
    args.X = mv.X;
    args.Xraw = mv.Xraw;
    args.x_stride = mv.x_stride;
    args.V = mv.V;
    args.v_cstride = mv.v_cstride;
    args.v_istride = mv.v_istride;
    args.src = mv.src;
    args.tma_mask = 0;
    args.part = mv.part;
    for (int d = 0; d < 3; ++d)
    {
        args.sel_lo[d] = mv.sel_lo[d];
        args.sel_hi[d] = mv.sel_hi[d];
    }
    static const bool no_tma = getenv("IBK_NO_TMA") != nullptr;
    static const bool dbg = getenv("IBK_DEBUG") != nullptr;
    static const int promo = getenv("IBK_TMA_PROMO_INTERP") ? atoi(getenv("IBK_TMA_PROMO_INTERP")) : 2;
    for (int a = 0; a < tp.ncomp; ++a)
        if (!no_tma && make_tensor_map(&maps.m[a], tp.comp[a], NDIM, S + 2, S, S, promo)) args.tma_mask |= (1u << a);
    if (dbg)
        fprintf(stderr, "[ibk] interp<%d,%d> ncomp=%d tma_mask=%x ntiles=%d n=(%d,%d,%d) pitch=%lld ptr=%p sizeof(tp)=%zu\n", NDIM, K,
                tp.ncomp, args.tma_mask, tp.nt[0] * tp.nt[1] * tp.nt[2], tp.comp[0].n[0], tp.comp[0].n[1], tp.comp[0].n[2],
                tp.comp[0].pitch, (void*)tp.comp[0].ptr, sizeof(TileParams));
    const int ntiles = tp.nt[0] * tp.nt[1] * tp.nt[2];
    if (ntiles <= 0) return cudaSuccess;
    auto go = [&](auto kfn, int nt) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
        {
            err = "cudaFuncSetAttribute(interp) failed";
            return e;
        }
        kfn<<<ntiles, nt, smem, L.stream>>>(tp, maps, args);
        L.launches++;
        return cudaGetLastError();
    };
    if constexpr (NDIM == 3 && M == 2 && KTraits<K>::W == 4)
    {
        // the conflict-free gather needs every component staged by TMA
        static const char* env = getenv("IBK_INTERP_ROT"); // 0 / 1 overrides the default
        static const bool rot = env ? atoi(env) != 0 : INTERP_ROT_DEFAULT;
        if (rot && args.tma_mask == (1u << tp.ncomp) - 1u) return go(interp_rot_kernel<K, INTERP_THREADS_WIDE>, INTERP_THREADS_WIDE);
    }
    return go(interp_tile_kernel<NDIM, K, INTERP_THREADS>, INTERP_THREADS);
}

template <int NDIM>
static cudaError_t launch_interp_k(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                                   std::string& err)
{
    switch (kernel)
    {
    case IBK_PIECEWISE_LINEAR:
        return launch_interp_t<NDIM, IBK_PIECEWISE_LINEAR>(L, tp, bins, mv, err);
    case IBK_IB_4:
        return launch_interp_t<NDIM, IBK_IB_4>(L, tp, bins, mv, err);
    case IBK_IB_6:
        return launch_interp_t<NDIM, IBK_IB_6>(L, tp, bins, mv, err);
    case IBK_BSPLINE_3:
        return launch_interp_t<NDIM, IBK_BSPLINE_3>(L, tp, bins, mv, err);
    case IBK_BSPLINE_4:
        return launch_interp_t<NDIM, IBK_BSPLINE_4>(L, tp, bins, mv, err);
    case IBK_IB_3:
        return launch_interp_t<NDIM, IBK_IB_3>(L, tp, bins, mv, err);
    case IBK_BSPLINE_5:
        return launch_interp_t<NDIM, IBK_BSPLINE_5>(L, tp, bins, mv, err);
    case IBK_BSPLINE_6:
        return launch_interp_t<NDIM, IBK_BSPLINE_6>(L, tp, bins, mv, err);
    case IBK_PIECEWISE_CUBIC:
        return launch_interp_t<NDIM, IBK_PIECEWISE_CUBIC>(L, tp, bins, mv, err);
    case IBK_IB_5:
        return launch_interp_t<NDIM, IBK_IB_5>(L, tp, bins, mv, err);
    case IBK_PIECEWISE_CONSTANT:
        return launch_interp_t<NDIM, IBK_PIECEWISE_CONSTANT>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_32:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_32>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_23:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_23>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_43:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_43>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_34:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_34>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_54:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_54>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_45:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_45>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_65:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_65>(L, tp, bins, mv, err);
    case IBK_COMPOSITE_BSPLINE_56:
        return launch_interp_t<NDIM, IBK_COMPOSITE_BSPLINE_56>(L, tp, bins, mv, err);
    case IBK_DISCONTINUOUS_LINEAR:
        return launch_interp_t<NDIM, IBK_DISCONTINUOUS_LINEAR>(L, tp, bins, mv, err);
    case IBK_IB_4_W8:
        return launch_interp_t<NDIM, IBK_IB_4_W8>(L, tp, bins, mv, err);
    default:
        err = "unknown kernel";
        return cudaErrorInvalidValue;
    }
}

cudaError_t launch_interp(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                          std::string& err)
{
    if (tp.ndim == 3) return launch_interp_k<3>(L, kernel, tp, bins, mv, err);
    return launch_interp_k<2>(L, kernel, tp, bins, mv, err);
}

} // namespace ibk
