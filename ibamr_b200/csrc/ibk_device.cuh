// ibk_device.cuh -- device-side building blocks shared by the sm_100a kernels:
// delta-kernel stencils (origin + 1-D weights), geometry structs, small PTX wrappers.
//
// Arithmetic follows the reference Fortran (ibtk/src/lagrangian/fortran/
// lagrangian_interaction3d.f.m4, lagrangian_delta.f.m4; line numbers at each function).
// Everything that decides an INTEGER (cell index, stencil origin, left/right choice) is written
// with explicit round-to-nearest intrinsics (__dsub_rn, __ddiv_rn, ...) so that nvcc cannot
// contract it into FMAs: those integers must be bit-identical to the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ibk.h"

namespace ibk
{
constexpr int BRICK = 4;       // cells per brick edge (binning granularity)
constexpr int TILE_BRICKS = 4; // bricks per marker-tile edge
constexpr int TILE = BRICK * TILE_BRICKS; // 16 cells / points per tile edge

// ---------------------------------------------------------------------------------------------
// kernel traits: W = points per dimension, M = reach (cells) of the stencil from the marker's cell
// ---------------------------------------------------------------------------------------------
template <int K>
struct KTraits;
template <>
struct KTraits<IBK_PIECEWISE_LINEAR>
{
    static constexpr int W = 2, M = 1;
};
template <>
struct KTraits<IBK_IB_4>
{
    static constexpr int W = 4, M = 2;
};
template <>
struct KTraits<IBK_IB_6>
{
    static constexpr int W = 6, M = 3;
};
template <>
struct KTraits<IBK_BSPLINE_3>
{
    static constexpr int W = 3, M = 2;
};
template <>
struct KTraits<IBK_BSPLINE_4>
{
    static constexpr int W = 4, M = 2;
};
// N4 kernels.  M counts from the marker's BINNING cell: for the side-shifted dimension the stencil's centre cell can
// be one above it, hence M = h + 1 for the centred rules (ib_3 h = 1, bspline_5 h = 2).
template <>
struct KTraits<IBK_IB_3>
{
    static constexpr int W = 3, M = 2;
};
template <>
struct KTraits<IBK_BSPLINE_5>
{
    static constexpr int W = 5, M = 3;
};
template <>
struct KTraits<IBK_BSPLINE_6>
{
    static constexpr int W = 6, M = 3;
};
template <>
struct KTraits<IBK_PIECEWISE_CUBIC>
{
    static constexpr int W = 4, M = 2;
};
template <>
struct KTraits<IBK_IB_5>
{
    static constexpr int W = 5, M = 3;
};
template <>
struct KTraits<IBK_PIECEWISE_CONSTANT>
{
    static constexpr int W = 1, M = 1;
};
// N4, second part: kernels whose 1-D function depends on whether the dimension is the component's axis
// (3d.f.m4:237-490, 3510-5460) and the 8-point ib_4.  Index rules are those of the wider member of the pair.
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_32>
{
    static constexpr int W = 3, M = 2;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_23>
{
    static constexpr int W = 3, M = 2;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_43>
{
    static constexpr int W = 4, M = 2;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_34>
{
    static constexpr int W = 4, M = 2;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_54>
{
    static constexpr int W = 5, M = 3;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_45>
{
    static constexpr int W = 5, M = 3;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_65>
{
    static constexpr int W = 6, M = 3;
};
template <>
struct KTraits<IBK_COMPOSITE_BSPLINE_56>
{
    static constexpr int W = 6, M = 3;
};
template <>
struct KTraits<IBK_DISCONTINUOUS_LINEAR>
{
    static constexpr int W = 2, M = 1; // off the axis: the centre cell with weight 1 and a second point with weight 0
};
template <>
struct KTraits<IBK_IB_4_W8>
{
    static constexpr int W = 8, M = 4;
};

// Fortran NINT (round half away from zero).
__device__ __forceinline__ int nint_f(double x)
{
    return (int)round(x);
}

// lagrangian_delta.f.m4:243-260
__device__ __forceinline__ double bspline_3_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 1.5;
    const double r2 = r * r;
    if (modx <= 0.5) return 0.5 * (-2.0 * r2 + 6.0 * r - 3.0);
    if (modx <= 1.5) return 0.5 * (r2 - 6.0 * r + 9.0);
    return 0.0;
}
// lagrangian_delta.f.m4:268-288
__device__ __forceinline__ double bspline_4_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 2.0;
    const double r2 = r * r;
    const double r3 = r2 * r;
    if (modx <= 1.0) return (1.0 / 6.0) * (3.0 * r3 - 24.0 * r2 + 60.0 * r - 44.0);
    if (modx <= 2.0) return (1.0 / 6.0) * (-r3 + 12.0 * r2 - 48.0 * r + 64.0);
    return 0.0;
}

// lagrangian_delta.f.m4:123-143 (sixth / third are the reference's truncated decimals)
__device__ __forceinline__ double ib_3_delta(double r)
{
    const double sixth = 0.16666666666667, third = 0.333333333333333;
    r = fabs(r);
    if (r < 0.5) return third * (1.0 + sqrt(1.0 - 3.0 * r * r));
    if (r < 1.5) return sixth * (5.0 - 3.0 * r - sqrt(1.0 - 3.0 * (1.0 - r) * (1.0 - r)));
    return 0.0;
}
// lagrangian_delta.f.m4:296-320
__device__ __forceinline__ double bspline_5_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 2.5;
    const double r2 = r * r, r3 = r2 * r, r4 = r3 * r;
    if (modx <= 0.5) return (1.0 / 24.0) * (6.0 * r4 - 60.0 * r3 + 210.0 * r2 - 300.0 * r + 155.0);
    if (modx <= 1.5) return (1.0 / 24.0) * (-4.0 * r4 + 60.0 * r3 - 330.0 * r2 + 780.0 * r - 655.0);
    if (modx <= 2.5) return (1.0 / 24.0) * (r4 - 20.0 * r3 + 150.0 * r2 - 500.0 * r + 625.0);
    return 0.0;
}
// lagrangian_delta.f.m4:328-352
__device__ __forceinline__ double bspline_6_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 3.0;
    const double r2 = r * r, r3 = r2 * r, r4 = r3 * r, r5 = r4 * r;
    if (modx <= 1.0) return (1.0 / 60.0) * (2193.0 - 3465.0 * r + 2130.0 * r2 - 630.0 * r3 + 90.0 * r4 - 5.0 * r5);
    if (modx <= 2.0) return (1.0 / 120.0) * (-10974.0 + 12270.0 * r - 5340.0 * r2 + 1140.0 * r3 - 120.0 * r4 + 5.0 * r5);
    if (modx <= 3.0) return (1.0 / 120.0) * (7776.0 - 6480.0 * r + 2160.0 * r2 - 360.0 * r3 + 30.0 * r4 - r5);
    return 0.0;
}
// lagrangian_delta.f.m4:74-93
__device__ __forceinline__ double piecewise_cubic_delta(double r)
{
    r = fabs(r);
    if (r < 1.0) return 1.0 - 0.5 * r - r * r + 0.5 * r * r * r;
    if (r < 2.0) return 1.0 - (11.0 / 6.0) * r + r * r - (1.0 / 6.0) * r * r * r;
    return 0.0;
}

// lagrangian_delta.f.m4:29-45
__device__ __forceinline__ double piecewise_linear_delta(double r)
{
    r = fabs(r);
    return (r < 1.0) ? 1.0 - r : 0.0;
}
// the B-spline of order `order` (2 = piecewise linear)
template <int order>
__device__ __forceinline__ double bspline_delta(double r)
{
    if constexpr (order == 2) return piecewise_linear_delta(r);
    if constexpr (order == 3) return bspline_3_delta(r);
    if constexpr (order == 4) return bspline_4_delta(r);
    if constexpr (order == 5) return bspline_5_delta(r);
    return bspline_6_delta(r);
}
// COMPOSITE_BSPLINE_<A><B>: order A along the component's axis, order B in the other dimensions
template <int K>
struct CompositeOrders
{
    static constexpr bool is = false;
    static constexpr int A = 0, B = 0;
};
#define IBK_COMPOSITE(AA, BB)                            \
    template <>                                          \
    struct CompositeOrders<IBK_COMPOSITE_BSPLINE_##AA##BB> \
    {                                                    \
        static constexpr bool is = true;                 \
        static constexpr int A = AA, B = BB;             \
    };
IBK_COMPOSITE(3, 2)
IBK_COMPOSITE(2, 3)
IBK_COMPOSITE(4, 3)
IBK_COMPOSITE(3, 4)
IBK_COMPOSITE(5, 4)
IBK_COMPOSITE(4, 5)
IBK_COMPOSITE(6, 5)
IBK_COMPOSITE(5, 6)
#undef IBK_COMPOSITE

// One dimension of a stencil.  Inputs: Xs = X + Xshift, Xraw = X (BSPLINE_4 quirk), x_lower and
// dx of the array, all exactly as the Fortran receives them.  Output: `lo` = first stencil index
// RELATIVE to the array's ilower (ic_lower - ilower) and W weights; weight j belongs to index
// lo + j.  Clipping to the ghost box is the caller's job: every in-scope kernel's weights depend
// only on the point's own index, so "clamp the bounds, then weigh" (3d.f.m4:2666-2678) equals
// "weigh, then skip the clipped points".
// on_axis: the dimension is the `axis` argument of the axis-dependent Fortran routines (ignored by the others).
template <int K>
__device__ __forceinline__ void stencil_1d(double Xs, double Xraw, double x_lower, double dx, int& lo, double* w, bool on_axis = false)
{
    // (X + Xshift - x_lower)/dx, 3d.f.m4:1265 / :2660 / :565
    const double t = __ddiv_rn(__dsub_rn(Xs, x_lower), dx);
    if constexpr (K == IBK_IB_4)
    {
        // 3d.f.m4:1266-1273
        lo = nint_f(t) - 2;
        const double r = __dsub_rn(t, __dadd_rn((double)(lo + 1), 0.5));
        const double q = sqrt(1.0 + 4.0 * r * (1.0 - r));
        w[0] = 0.125 * (3.0 - 2.0 * r - q);
        w[1] = 0.125 * (3.0 - 2.0 * r + q);
        w[2] = 0.125 * (1.0 + 2.0 * r + q);
        w[3] = 0.125 * (1.0 + 2.0 * r - q);
    }
    else if constexpr (K == IBK_IB_6)
    {
        // 3d.f.m4:2220, 2244-2272
        // K = (59/60)*(1 - sqrt(1 - 3220/3481)) evaluated in double precision
        const double Kc = 0x1.6d9b402672048p-1; // 0.714075092976608
        lo = nint_f(t) - 3;
        const double r = (1.0 - t) + ((double)(lo + 2) + 0.5);
        const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, r6 = r4 * r2;
        const double alpha = 28.0;
        const double beta = (9.0 / 4.0) - (3.0 / 2.0) * (Kc + r2) + ((22.0 / 3.0) - 7.0 * Kc) * r - (7.0 / 3.0) * r3;
        const double gamma = (1.0 / 4.0) * (((161.0 / 36.0) - (59.0 / 6.0) * Kc + 5.0 * (Kc * Kc)) * (1.0 / 2.0) * r2 +
                                            (-(109.0 / 24.0) + 5.0 * Kc) * (1.0 / 3.0) * r4 + (5.0 / 18.0) * r6);
        const double discr = beta * beta - 4.0 * alpha * gamma;
        // sign(1, 3/2 - K) = +1 since K ~ 0.714
        const double pm3 = (-beta + sqrt(discr)) / (2.0 * alpha);
        w[0] = pm3;
        w[1] = -3.0 * pm3 - (1.0 / 16.0) + (1.0 / 8.0) * (Kc + r2) + (1.0 / 12.0) * (3.0 * Kc - 1.0) * r +
               (1.0 / 12.0) * r3;
        w[2] = 2.0 * pm3 + (1.0 / 4.0) + (1.0 / 6.0) * (4.0 - 3.0 * Kc) * r - (1.0 / 6.0) * r3;
        w[3] = 2.0 * pm3 + (5.0 / 8.0) - (1.0 / 4.0) * (Kc + r2);
        w[4] = -3.0 * pm3 + (1.0 / 4.0) - (1.0 / 6.0) * (4.0 - 3.0 * Kc) * r + (1.0 / 6.0) * r3;
        w[5] = pm3 - (1.0 / 16.0) + (1.0 / 8.0) * (Kc + r2) - (1.0 / 12.0) * (3.0 * Kc - 1.0) * r - (1.0 / 12.0) * r3;
    }
    else if constexpr (K == IBK_BSPLINE_3)
    {
        // 3d.f.m4:2659-2678: centre cell floor(t); weight = delta((Xs - X_cell)/dx)
        const int c = (int)floor(t);
        lo = c - 1;
#pragma unroll
        for (int j = 0; j < 3; ++j)
        {
            const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)(lo + j), 0.5), dx));
            w[j] = bspline_3_delta(__ddiv_rn(__dsub_rn(Xs, X_cell), dx));
        }
    }
    else if constexpr (K == IBK_BSPLINE_4)
    {
        // 3d.f.m4:2882-2908; the side choice compares the UNSHIFTED X with X_cell (:2891)
        const int c = (int)floor(t);
        const double X_cell_c = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)c, 0.5), dx));
        lo = (Xraw < X_cell_c) ? c - 2 : c - 1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)(lo + j), 0.5), dx));
            w[j] = bspline_4_delta(__ddiv_rn(__dsub_rn(Xs, X_cell), dx));
        }
    }
    else if constexpr (K == IBK_IB_5)
    {
        // lagrangian_ib_5_interp3d: centre cell floor(t), points c-2..c+2, r = (Xs - X_cell(c))/dx
        const double Kc = (38.0 - 8.306623862918075) / 60.0; // (38 - sqrt(69))/60, sqrt(69) correctly rounded
        const int c = (int)floor(t);
        lo = c - 2;
        const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)c, 0.5), dx));
        const double r = __ddiv_rn(__dsub_rn(Xs, X_cell), dx);
        const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, r6 = r4 * r2;
        const double phi = (136.0 - 40.0 * Kc - 40.0 * r2 +
                            1.4142135623730951 * sqrt(3123.0 - 6840.0 * Kc + 3600.0 * (Kc * Kc) - 12440.0 * r2 + 25680.0 * Kc * r2 -
                                                      12600.0 * (Kc * Kc) * r2 + 8080.0 * r4 - 8400.0 * Kc * r4 - 1400.0 * r6)) /
                           280.0;
        w[0] = (1.0 / 12.0) * (-2.0 + 2.0 * phi + 2.0 * Kc + r - 3.0 * Kc * r + 2.0 * r2 - r3);
        w[1] = (1.0 / 6.0) * (4.0 - 4.0 * phi - Kc - 4.0 * r + 3.0 * Kc * r - r2 + r3);
        w[2] = phi;
        w[3] = (1.0 / 6.0) * (4.0 - 4.0 * phi - Kc + 4.0 * r - 3.0 * Kc * r - r2 - r3);
        w[4] = (1.0 / 12.0) * (-2.0 + 2.0 * phi + 2.0 * Kc - r + 3.0 * Kc * r + 2.0 * r2 + r3);
    }
    else if constexpr (K == IBK_PIECEWISE_CONSTANT)
    {
        // lagrangian_piecewise_constant_interp3d: the cell NINT(t - 0.5), weight 1
        lo = nint_f(__dsub_rn(t, 0.5));
        w[0] = 1.0;
    }
    else if constexpr (K == IBK_IB_3 || K == IBK_BSPLINE_5)
    {
        // centred rule [c - h, c + h] (3d.f.m4: ib_3 :1038-1066, bspline_5), weight = delta((Xs - X_cell)/dx)
        constexpr int h = (K == IBK_IB_3) ? 1 : 2;
        const int c = (int)floor(t);
        lo = c - h;
#pragma unroll
        for (int j = 0; j < 2 * h + 1; ++j)
        {
            const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)(lo + j), 0.5), dx));
            const double r = __ddiv_rn(__dsub_rn(Xs, X_cell), dx);
            w[j] = (K == IBK_IB_3) ? ib_3_delta(r) : bspline_5_delta(r);
        }
    }
    else if constexpr (K == IBK_BSPLINE_6 || K == IBK_PIECEWISE_CUBIC)
    {
        // sided rule (3d.f.m4: bspline_6, piecewise_cubic): the UNSHIFTED X against X_cell(c) picks [c - h, c + h - 1]
        // or [c - h + 1, c + h]
        constexpr int h = (K == IBK_BSPLINE_6) ? 3 : 2;
        const int c = (int)floor(t);
        const double X_cell_c = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)c, 0.5), dx));
        lo = (Xraw < X_cell_c) ? c - h : c - h + 1;
#pragma unroll
        for (int j = 0; j < 2 * h; ++j)
        {
            const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)(lo + j), 0.5), dx));
            const double r = __ddiv_rn(__dsub_rn(Xs, X_cell), dx);
            w[j] = (K == IBK_BSPLINE_6) ? bspline_6_delta(r) : piecewise_cubic_delta(r);
        }
    }
    else if constexpr (CompositeOrders<K>::is)
    {
        // composite B-splines (3d.f.m4:3510-5460): centred rule [c - h, c + h] for the odd widths, sided rule (the
        // UNSHIFTED X against X_cell(c)) for the even ones; weight = delta_A on the axis, delta_B off it
        constexpr int W = KTraits<K>::W;
        constexpr int h = W / 2;
        const int c = (int)floor(t);
        if constexpr (W % 2 == 1)
            lo = c - h;
        else
        {
            const double X_cell_c = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)c, 0.5), dx));
            lo = (Xraw < X_cell_c) ? c - h : c - h + 1;
        }
#pragma unroll
        for (int j = 0; j < W; ++j)
        {
            const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)(lo + j), 0.5), dx));
            const double r = __ddiv_rn(__dsub_rn(Xs, X_cell), dx);
            w[j] = on_axis ? bspline_delta<CompositeOrders<K>::A>(r) : bspline_delta<CompositeOrders<K>::B>(r);
        }
    }
    else if constexpr (K == IBK_IB_4_W8)
    {
        // 3d.f.m4:1545-1560: first point NINT(t) - 4; odd points from r = (t - (lo + 3 + 1/2))/2, even ones from r + 1/2
        lo = nint_f(t) - 4;
        double r = 0.5 * __dsub_rn(t, __dadd_rn((double)(lo + 3), 0.5));
        double q = sqrt(1.0 + 4.0 * r * (1.0 - r));
        w[1] = 0.0625 * (3.0 - 2.0 * r - q);
        w[3] = 0.0625 * (3.0 - 2.0 * r + q);
        w[5] = 0.0625 * (1.0 + 2.0 * r + q);
        w[7] = 0.0625 * (1.0 + 2.0 * r - q);
        r = r + 0.5;
        q = sqrt(1.0 + 4.0 * r * (1.0 - r));
        w[0] = 0.0625 * (3.0 - 2.0 * r - q);
        w[2] = 0.0625 * (3.0 - 2.0 * r + q);
        w[4] = 0.0625 * (1.0 + 2.0 * r + q);
        w[6] = 0.0625 * (1.0 + 2.0 * r - q);
    }
    else
    {
        // PIECEWISE_LINEAR, 3d.f.m4:563-579; DISCONTINUOUS_LINEAR (3d.f.m4:296-321) is the same along the axis and the
        // centre cell alone (weight 1; the second point carries weight 0) in the other dimensions
        const int c = nint_f(__dsub_rn(t, 0.5));
        if (K == IBK_DISCONTINUOUS_LINEAR && !on_axis)
        {
            lo = c;
            w[0] = 1.0;
            w[1] = 0.0;
            return;
        }
        const double X_cell = __dadd_rn(x_lower, __dmul_rn(__dadd_rn((double)c, 0.5), dx));
        if (Xs < X_cell)
        {
            lo = c - 1;
            w[0] = __ddiv_rn(__dsub_rn(X_cell, Xs), dx);
        }
        else
        {
            lo = c;
            w[0] = 1.0 + __ddiv_rn(__dsub_rn(X_cell, Xs), dx);
        }
        w[1] = 1.0 - w[0];
    }
}

// IndexUtilities::getCellIndex, one dimension (ibtk/include/ibtk/private/IndexUtilities-inl.h:62-73).
__device__ __forceinline__ int cell_index_1d(double X, double x_lower, double x_upper, double dx, int ilower, int iupper)
{
    const double dX_lower = __dsub_rn(X, x_lower);
    const double dX_upper = __dsub_rn(X, x_upper);
    if (fabs(dX_lower) <= fabs(dX_upper)) return ilower + (int)floor(__ddiv_rn(dX_lower, dx));
    return iupper + (int)floor(__ddiv_rn(dX_upper, dx)) + 1;
}

// ---------------------------------------------------------------------------------------------
// geometry handed to the tile kernels
// ---------------------------------------------------------------------------------------------
// One scalar array (one SideData axis, or one depth slice of a CellData) in "pp" coordinates:
// pp = array index - (patch_lower - G), the same origin the binning uses for cells.
struct CompGeom
{
    double* ptr;           // element (0,0,0) of the array (first ghost)
    long long pitch;       // elements between consecutive rows (>= n[0])
    int n[3];              // extents incl. ghosts
    int pp0[3];            // pp coordinate of element 0  (= G - nugc)
    int var[3];            // which x_lower variant each dimension uses (see TileParams::xl)
    int vcol;              // which marker value column this component reads / writes
    int axis;              // the `axis` argument of the axis-dependent kernels (SideData / EdgeData: the component)
};

struct TileParams
{
    int ndim;
    int ncomp;
    CompGeom comp[IBK_MAX_COMP];
    double dx[3];
    double xl[3][2];       // x_lower variants per dimension, RELATIVE indexing: the stencil's `lo`
                           // is relative to the patch lower index; variant 0 = cell-centred,
                           // 1 = shifted by -dx/2 (LEInteractor.cpp:2464)
    int nvar[3];
    int G;                 // binning margin: pp = (index - patch_lower) + G
    int nb[3];             // bricks per dimension (padded to a multiple of TILE_BRICKS)
    int nt[3];             // marker tiles per dimension
    int brick_base;        // first brick id of this patch in brick_start[]
    int ot_lo[3];          // first output tile index per dimension (spread)
    int ot_n[3];           // number of output tiles per dimension (spread)
    double inv_vol;        // 1 / (dx0*dx1[*dx2])  (spread scale, 3d.f.m4:1439)
};

// Hierarchical brick id: tiles (tz,ty,tx) row-major, then brick-in-tile (bz,by,bx).
__host__ __device__ __forceinline__ int brick_id_3d(int bx, int by, int bz, const int* nt)
{
    const int tx = bx >> 2, ty = by >> 2, tz = bz >> 2;
    return (((tz * nt[1] + ty) * nt[0] + tx) << 6) | ((bz & 3) << 4) | ((by & 3) << 2) | (bx & 3);
}
__host__ __device__ __forceinline__ int brick_id_2d(int bx, int by, const int* nt)
{
    const int tx = bx >> 2, ty = by >> 2;
    return ((ty * nt[0] + tx) << 4) | ((by & 3) << 2) | (bx & 3);
}

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA (cp.async.bulk.tensor), sm_90+/sm_100a
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    while (!mbar_try_wait(bar, phase))
    {
    }
}
// generic-proxy writes to shared memory -> later async-proxy (TMA) accesses of the same bytes
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// shared memory -> global box store (clipped to the tensor), completion tracked by a bulk async-group
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}


// Walks a slab row by row: a warp takes one row (contiguous in x) and calls f(I1, I2, x_first, x_last) with x advancing by
// the lane stride 32 (coalesced 256-byte accesses, the row's index arithmetic done once per row); thin slabs (the x ghost
// layers, a few elements per row) are walked with 32 / e0 rows per warp, one element per lane.
template <class F>
__device__ __forceinline__ void slab_rows(const int* lo, const int* hi, unsigned bx, unsigned nbx, F f)
{
    const int e0 = hi[0] - lo[0] + 1, e1 = hi[1] - lo[1] + 1, e2 = hi[2] - lo[2] + 1;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const long long nrows = (long long)e1 * e2;
    if (e0 >= 16)
    {
        // work item = (row, chunk of 128 elements): a slab of a few long rows (the z ghost layers: 3 x 518 rows of 519) still
        // spreads over thousands of warps, each with four independent elements per lane in flight
        constexpr int CHUNK = 128;
        const int nch = (e0 + CHUNK - 1) / CHUNK;
        const long long nitems = nrows * nch;
        for (long long it = (long long)bx * wpb + warp; it < nitems; it += (long long)nbx * wpb)
        {
            const long long row = it / nch;
            const int ch = (int)(it - row * nch);
            const int k = (int)(row / e1);
            const int xs = lo[0] + ch * CHUNK;
            f(lo[1] + (int)(row - (long long)k * e1), lo[2] + k, xs + lane, min(hi[0], xs + CHUNK - 1));
        }
    }
    else
    {
        const int rpw = 32 / e0, r = lane / e0, x = lo[0] + lane - r * e0;
        for (long long g = (long long)bx * wpb + warp; g * rpw < nrows; g += (long long)nbx * wpb)
        {
            const long long row = g * rpw + r;
            if (r >= rpw || row >= nrows) continue;
            const int k = (int)(row / e1);
            f(lo[1] + (int)(row - (long long)k * e1), lo[2] + k, x, x);
        }
    }
}

} // namespace ibk
