/*
 * oracle/le_kernels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, no dependencies) of the arithmetic of IBAMR's
 * Lagrangian-Eulerian interaction Fortran routines for the five in-scope
 * kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call into this file; the product path
 * (ibamr_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * the reference's own golden files (copied fixtures under tests/golden/):
 *   tests/interpolate/interpolate_01_{2d,3d}.{ib_4,ib_6,bspline_3,bspline_4,
 *   piecewise_linear}.output, tests/IBTK/ghost_accumulation_01_*spread*.output,
 *   tests/IBTK/index_utilities_{2d,3d}.output.
 *
 * Reference files followed (all under /root/reference/ibtk/src/lagrangian/fortran):
 *   lagrangian_interaction3d.f.m4:24-86     fixed-width inner loops (bspline)
 *   lagrangian_interaction3d.f.m4:493-730   piecewise_linear interp/spread
 *   lagrangian_interaction3d.f.m4:1203-1475 ib_4 interp/spread
 *   lagrangian_interaction3d.f.m4:2178-2582 ib_6 interp/spread
 *   lagrangian_interaction3d.f.m4:2591-2805 bspline_3 interp/spread
 *   lagrangian_interaction3d.f.m4:2814-3041 bspline_4 interp/spread
 *   lagrangian_interaction2d.f.m4:469,587,1152,1273,1971,2136,2300,2410,2519,2634
 *   lagrangian_delta.f.m4:243-288           bspline_3 / bspline_4 1-D functions
 *
 * Structure here is NOT the reference's: the Fortran has one hand-unrolled
 * subroutine per (kernel, op, dim).  This file factors every routine into
 *   (1) a per-dimension "stencil" (loop bounds + 1-D weights), and
 *   (2) one of two accumulation styles:
 *       TENSOR  (ib_4, ib_6): w3 = w0*(w1*wz), wz = w2 [/ (dx0*dx1*dx2) for spread]
 *       INLINE  (bspline_3/4, piecewise_linear): ((w0*w1)*w2)*u, spread divides
 *               the whole term by (dx0*dx1*dx2)
 * keeping the reference's evaluation order so results agree to rounding.
 * Build with -ffp-contract=off (no FMA contraction), see oracle/Makefile.
 *
 * Array conventions (Fortran, restated):
 *   u(ilower0-g0:iupper0+g0, ilower1-g1:iupper1+g1[, ilower2-g2:iupper2+g2], 0:depth-1)
 *   X(0:NDIM-1, 0:*), V(0:depth-1, 0:*), Xshift(0:NDIM-1, 0:nindices-1), indices(0:nindices-1)
 */
#include <math.h>
#include <stddef.h>

#define MAXW 8

enum
{
    K_PIECEWISE_LINEAR = 0,
    K_IB_4 = 1,
    K_IB_6 = 2,
    K_BSPLINE_3 = 3,
    K_BSPLINE_4 = 4,
    /* N4: four more of the reference's kernels, all of the "delta function per stencil point" kind */
    K_IB_3 = 5,
    K_BSPLINE_5 = 6,
    K_BSPLINE_6 = 7,
    K_PIECEWISE_CUBIC = 8,
    K_IB_5 = 9,
    K_PIECEWISE_CONSTANT = 10,
    /* N4, second part: the kernels whose 1-D function depends on whether the dimension is the component's `axis`
     * (the Fortran routines take an extra `axis` argument, 3d.f.m4:237, 3510-5460), and the broadened ib_4.
     * The axis travels in bits 8.. of the `kernel` argument of le_oracle_interp / le_oracle_spread. */
    K_COMPOSITE_BSPLINE_32 = 11,
    K_COMPOSITE_BSPLINE_23 = 12,
    K_COMPOSITE_BSPLINE_43 = 13,
    K_COMPOSITE_BSPLINE_34 = 14,
    K_COMPOSITE_BSPLINE_54 = 15,
    K_COMPOSITE_BSPLINE_45 = 16,
    K_COMPOSITE_BSPLINE_65 = 17,
    K_COMPOSITE_BSPLINE_56 = 18,
    K_DISCONTINUOUS_LINEAR = 19,
    K_IB_4_W8 = 20
};

typedef struct
{
    int lo, hi;     /* loop bounds in array index space (already clipped to the ghost box) */
    int wbase;      /* weight j belongs to array index wbase + j                            */
    double w[MAXW]; /* 1-D weights                                                          */
} stencil1d;

/* Fortran NINT: round half away from zero. */
static inline int nint_f(double x)
{
    return (int)lround(x);
}

/* lagrangian_delta.f.m4:243-260 */
static double bspline_3_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 1.5;
    const double r2 = r * r;
    if (modx <= 0.5) return 0.5 * (-2.0 * r2 + 6.0 * r - 3.0);
    if (modx <= 1.5) return 0.5 * (r2 - 6.0 * r + 9.0);
    return 0.0;
}

/* lagrangian_delta.f.m4:268-288 */
static double bspline_4_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 2.0;
    const double r2 = r * r;
    const double r3 = r2 * r;
    if (modx <= 1.0) return (1.0 / 6.0) * (3.0 * r3 - 24.0 * r2 + 60.0 * r - 44.0);
    if (modx <= 2.0) return (1.0 / 6.0) * (-r3 + 12.0 * r2 - 48.0 * r + 64.0);
    return 0.0;
}

/* 3d.f.m4:1265-1273, 1309-1310 (one dimension of ib_4). */
static void stencil_ib_4(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const double X_o_dx = (Xs - x_lower) / dx;
    const int ic_lower = nint_f(X_o_dx) + ilower - 2;
    const int ic_upper = ic_lower + 3;
    const double r = X_o_dx - ((ic_lower + 1 - ilower) + 0.5);
    const double q = sqrt(1.0 + 4.0 * r * (1.0 - r));
    s->w[0] = 0.125 * (3.0 - 2.0 * r - q);
    s->w[1] = 0.125 * (3.0 - 2.0 * r + q);
    s->w[2] = 0.125 * (1.0 + 2.0 * r + q);
    s->w[3] = 0.125 * (1.0 + 2.0 * r - q);
    const int ig_lower = ilower - g, ig_upper = iupper + g;
    const int istart = (ig_lower - ic_lower > 0) ? ig_lower - ic_lower : 0;
    const int istop = 3 - ((ic_upper - ig_upper > 0) ? ic_upper - ig_upper : 0);
    s->wbase = ic_lower;
    s->lo = ic_lower + istart;
    s->hi = ic_lower + istop;
}

/* 3d.f.m4:2220, 2243-2272, 2350-2351 (one dimension of ib_6). */
static void stencil_ib_6(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const double K = (59.0 / 60.0) * (1.0 - sqrt(1.0 - (3220.0 / 3481.0)));
    const double X_o_dx = (Xs - x_lower) / dx;
    const int ic_lower = nint_f(X_o_dx) + ilower - 3;
    const int ic_upper = ic_lower + 5;
    const double r = 1.0 - X_o_dx + ((ic_lower + 2 - ilower) + 0.5);
    const double r2 = r * r;
    const double r3 = r2 * r;
    const double r4 = r2 * r2;
    const double r6 = r4 * r2;
    const double alpha = 28.0;
    const double beta = (9.0 / 4.0) - (3.0 / 2.0) * (K + r2) + ((22.0 / 3.0) - 7.0 * K) * r - (7.0 / 3.0) * r3;
    const double gamma = (1.0 / 4.0) * (((161.0 / 36.0) - (59.0 / 6.0) * K + 5.0 * (K * K)) * (1.0 / 2.0) * r2 +
                                        (-(109.0 / 24.0) + 5.0 * K) * (1.0 / 3.0) * r4 + (5.0 / 18.0) * r6);
    const double discr = beta * beta - 4.0 * alpha * gamma;
    const double sgn = ((3.0 / 2.0) - K) >= 0.0 ? 1.0 : -1.0;
    const double pm3 = (-beta + sgn * sqrt(discr)) / (2.0 * alpha);
    const double pm2 =
        -3.0 * pm3 - (1.0 / 16.0) + (1.0 / 8.0) * (K + r2) + (1.0 / 12.0) * (3.0 * K - 1.0) * r + (1.0 / 12.0) * r3;
    const double pm1 = 2.0 * pm3 + (1.0 / 4.0) + (1.0 / 6.0) * (4.0 - 3.0 * K) * r - (1.0 / 6.0) * r3;
    const double p = 2.0 * pm3 + (5.0 / 8.0) - (1.0 / 4.0) * (K + r2);
    const double pp1 = -3.0 * pm3 + (1.0 / 4.0) - (1.0 / 6.0) * (4.0 - 3.0 * K) * r + (1.0 / 6.0) * r3;
    const double pp2 =
        pm3 - (1.0 / 16.0) + (1.0 / 8.0) * (K + r2) - (1.0 / 12.0) * (3.0 * K - 1.0) * r - (1.0 / 12.0) * r3;
    s->w[0] = pm3;
    s->w[1] = pm2;
    s->w[2] = pm1;
    s->w[3] = p;
    s->w[4] = pp1;
    s->w[5] = pp2;
    const int ig_lower = ilower - g, ig_upper = iupper + g;
    const int istart = (ig_lower - ic_lower > 0) ? ig_lower - ic_lower : 0;
    const int istop = 5 - ((ic_upper - ig_upper > 0) ? ic_upper - ig_upper : 0);
    s->wbase = ic_lower;
    s->lo = ic_lower + istart;
    s->hi = ic_lower + istop;
}

/* lagrangian_delta.f.m4:123-143 (the constants sixth / third are the reference's truncated decimals) */
static double ib_3_delta(double r)
{
    const double sixth = 0.16666666666667, third = 0.333333333333333;
    if (r < 0.0) r = -r;
    if (r < 0.5) return third * (1.0 + sqrt(1.0 - 3.0 * r * r));
    if (r < 1.5) return sixth * (5.0 - 3.0 * r - sqrt(1.0 - 3.0 * (1.0 - r) * (1.0 - r)));
    return 0.0;
}
/* lagrangian_delta.f.m4:296-320 */
static double bspline_5_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 2.5;
    const double r2 = r * r, r3 = r2 * r, r4 = r3 * r;
    if (modx <= 0.5) return (1.0 / 24.0) * (6.0 * r4 - 60.0 * r3 + 210.0 * r2 - 300.0 * r + 155.0);
    if (modx <= 1.5) return (1.0 / 24.0) * (-4.0 * r4 + 60.0 * r3 - 330.0 * r2 + 780.0 * r - 655.0);
    if (modx <= 2.5) return (1.0 / 24.0) * (r4 - 20.0 * r3 + 150.0 * r2 - 500.0 * r + 625.0);
    return 0.0;
}
/* lagrangian_delta.f.m4:328-352 */
static double bspline_6_delta(double x)
{
    const double modx = fabs(x);
    const double r = modx + 3.0;
    const double r2 = r * r, r3 = r2 * r, r4 = r3 * r, r5 = r4 * r;
    if (modx <= 1.0) return (1.0 / 60.0) * (2193.0 - 3465.0 * r + 2130.0 * r2 - 630.0 * r3 + 90.0 * r4 - 5.0 * r5);
    if (modx <= 2.0) return (1.0 / 120.0) * (-10974.0 + 12270.0 * r - 5340.0 * r2 + 1140.0 * r3 - 120.0 * r4 + 5.0 * r5);
    if (modx <= 3.0) return (1.0 / 120.0) * (7776.0 - 6480.0 * r + 2160.0 * r2 - 360.0 * r3 + 30.0 * r4 - r5);
    return 0.0;
}
/* lagrangian_delta.f.m4:74-93 */
static double piecewise_cubic_delta(double r)
{
    if (r < 0.0) r = -r;
    if (r < 1.0) return 1.0 - 0.5 * r - r * r + 0.5 * r * r * r;
    if (r < 2.0) return 1.0 - (11.0 / 6.0) * r + r * r - (1.0 / 6.0) * r * r * r;
    return 0.0;
}

/* The four N4 kernels share two index rules (lagrangian_interaction3d.f.m4):
 *   centred, odd width 2h+1 (ib_3 :1038-1066 h = 1, bspline_5 h = 2): [c - h, c + h], c = floor((X+Xshift-x_lower)/dx)
 *   sided, even width 2h (piecewise_cubic h = 2, bspline_6 h = 3): unshifted X < X_cell(c) ? [c - h, c + h - 1] : [c - h + 1, c + h]
 * both clamped to the ghost box, weights indexed from the CLAMPED lower bound, weight = delta((X+Xshift-X_cell)/dx). */
static void stencil_delta(double (*delta)(double), int h, int sided, double Xs, double Xraw, double x_lower, double dx, int ilower,
                          int iupper, int g, stencil1d* s)
{
    const int ic_center = (int)floor((Xs - x_lower) / dx) + ilower;
    int ic_lower, ic_upper;
    if (!sided)
    {
        ic_lower = ic_center - h;
        ic_upper = ic_center + h;
    }
    else
    {
        const double X_cell_c = x_lower + ((double)(ic_center - ilower) + 0.5) * dx;
        if (Xraw < X_cell_c)
        {
            ic_lower = ic_center - h;
            ic_upper = ic_center + h - 1;
        }
        else
        {
            ic_lower = ic_center - h + 1;
            ic_upper = ic_center + h;
        }
    }
    if (ic_lower < ilower - g) ic_lower = ilower - g;
    if (ic_upper > iupper + g) ic_upper = iupper + g;
    for (int ic = ic_lower; ic <= ic_upper; ++ic)
    {
        const double X_cell = x_lower + ((double)(ic - ilower) + 0.5) * dx;
        s->w[ic - ic_lower] = delta((Xs - X_cell) / dx);
    }
    s->wbase = ic_lower;
    s->lo = ic_lower;
    s->hi = ic_upper;
}

/* ib_5 (lagrangian_interaction3d.f.m4, lagrangian_ib_5_interp3d): centre cell floor(...), five points c-2..c+2,
 * closed-form weights from r = (X - X_cell(c))/dx with K = (38 - sqrt(69))/60; loop bounds clipped to the ghost box,
 * weights indexed from the UNCLIPPED lower bound (istart/istop), products formed as w0*w1*w2*u. */
static void stencil_ib_5(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const double K = (38.0 - sqrt(69.0)) / 60.0;
    const int ic_center = (int)floor((Xs - x_lower) / dx) + ilower;
    const int ic_lower = ic_center - 2, ic_upper = ic_center + 2;
    const double X_cell = x_lower + ((double)(ic_center - ilower) + 0.5) * dx;
    const double r = (Xs - X_cell) / dx;
    const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, r6 = r4 * r2;
    const double phi = (136.0 - 40.0 * K - 40.0 * r2 +
                        sqrt(2.0) * sqrt(3123.0 - 6840.0 * K + 3600.0 * (K * K) - 12440.0 * r2 + 25680.0 * K * r2 -
                                         12600.0 * (K * K) * r2 + 8080.0 * r4 - 8400.0 * K * r4 - 1400.0 * r6)) /
                       280.0;
    s->w[0] = (1.0 / 12.0) * (-2.0 + 2.0 * phi + 2.0 * K + r - 3.0 * K * r + 2.0 * r2 - r3);
    s->w[1] = (1.0 / 6.0) * (4.0 - 4.0 * phi - K - 4.0 * r + 3.0 * K * r - r2 + r3);
    s->w[2] = phi;
    s->w[3] = (1.0 / 6.0) * (4.0 - 4.0 * phi - K + 4.0 * r - 3.0 * K * r - r2 - r3);
    s->w[4] = (1.0 / 12.0) * (-2.0 + 2.0 * phi + 2.0 * K - r + 3.0 * K * r + 2.0 * r2 + r3);
    const int ig_lower = ilower - g, ig_upper = iupper + g;
    const int istart = (ig_lower - ic_lower > 0) ? ig_lower - ic_lower : 0;
    const int istop = 4 - ((ic_upper - ig_upper > 0) ? ic_upper - ig_upper : 0);
    s->wbase = ic_lower;
    s->lo = ic_lower + istart;
    s->hi = ic_lower + istop;
}

/* piecewise_constant (lagrangian_piecewise_constant_interp3d): the one cell ilower + NINT((X-x_lower)/dx - 0.5), weight 1.
 * The Fortran does not clip (a point outside the ghost box would index out of the array); here such a point is skipped. */
static void stencil_piecewise_constant(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const int ic = ilower + nint_f((Xs - x_lower) / dx - 0.5);
    s->w[0] = 1.0;
    s->wbase = ic;
    s->lo = (ic >= ilower - g) ? ic : ic + 1; /* lo > hi: nothing */
    s->hi = (ic <= iupper + g) ? ic : ic - 1;
    if (ic < ilower - g) s->hi = ic - 1, s->lo = ic;
}

/* 3d.f.m4:2659-2678: weights are indexed from the CLAMPED lower bound. */
static void stencil_bspline_3(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const int ic_center = (int)floor((Xs - x_lower) / dx) + ilower;
    int ic_lower = ic_center - 1;
    int ic_upper = ic_center + 1;
    if (ic_lower < ilower - g) ic_lower = ilower - g;
    if (ic_upper > iupper + g) ic_upper = iupper + g;
    for (int ic = ic_lower; ic <= ic_upper; ++ic)
    {
        const double X_cell = x_lower + ((double)(ic - ilower) + 0.5) * dx;
        s->w[ic - ic_lower] = bspline_3_delta((Xs - X_cell) / dx);
    }
    s->wbase = ic_lower;
    s->lo = ic_lower;
    s->hi = ic_upper;
}

/* 3d.f.m4:2882-2908: the left/right choice compares the UNSHIFTED X with X_cell (:2891). */
static void
stencil_bspline_4(double Xs, double Xraw, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const int ic_center = (int)floor((Xs - x_lower) / dx) + ilower;
    const double X_cell_c = x_lower + ((double)(ic_center - ilower) + 0.5) * dx;
    int ic_lower, ic_upper;
    if (Xraw < X_cell_c)
    {
        ic_lower = ic_center - 2;
        ic_upper = ic_center + 1;
    }
    else
    {
        ic_lower = ic_center - 1;
        ic_upper = ic_center + 2;
    }
    if (ic_lower < ilower - g) ic_lower = ilower - g;
    if (ic_upper > iupper + g) ic_upper = iupper + g;
    for (int ic = ic_lower; ic <= ic_upper; ++ic)
    {
        const double X_cell = x_lower + ((double)(ic - ilower) + 0.5) * dx;
        s->w[ic - ic_lower] = bspline_4_delta((Xs - X_cell) / dx);
    }
    s->wbase = ic_lower;
    s->lo = ic_lower;
    s->hi = ic_upper;
}

/* 3d.f.m4:563-583: trimmed loop bounds, weights indexed from the UNTRIMMED lower bound. */
static void stencil_piecewise_linear(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const int ic_center = ilower + nint_f((Xs - x_lower) / dx - 0.5);
    const double X_cell = x_lower + ((double)(ic_center - ilower) + 0.5) * dx;
    int ic_lower, ic_upper;
    if (Xs < X_cell)
    {
        ic_lower = ic_center - 1;
        ic_upper = ic_center;
        s->w[0] = (X_cell - Xs) / dx;
        s->w[1] = 1.0 - s->w[0];
    }
    else
    {
        ic_lower = ic_center;
        ic_upper = ic_center + 1;
        s->w[0] = 1.0 + (X_cell - Xs) / dx;
        s->w[1] = 1.0 - s->w[0];
    }
    s->wbase = ic_lower;
    s->lo = (ic_lower > ilower - g) ? ic_lower : ilower - g;
    s->hi = (ic_upper < iupper + g) ? ic_upper : iupper + g;
}

/* lagrangian_delta.f.m4:29-45 */
static double piecewise_linear_delta(double r)
{
    if (r < 0.0) r = -r;
    return (r < 1.0) ? 1.0 - r : 0.0;
}

/* discontinuous_linear (3d.f.m4:296-336): centre cell NINT(t - 1/2); along `axis` the two-point hat of piecewise_linear,
 * in the other dimensions the centre cell alone with weight 1; trimmed bounds, weights indexed from the untrimmed lower
 * bound, products formed inline as w0*w1*w2*u. */
static void stencil_discontinuous_linear(int on_axis, double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    if (on_axis)
    {
        stencil_piecewise_linear(Xs, x_lower, dx, ilower, iupper, g, s);
        return;
    }
    const int ic_center = ilower + nint_f((Xs - x_lower) / dx - 0.5);
    s->w[0] = 1.0;
    s->wbase = ic_center;
    s->lo = (ic_center > ilower - g) ? ic_center : ilower - g;
    s->hi = (ic_center < iupper + g) ? ic_center : iupper + g;
}

/* ib_4_w8 (3d.f.m4:1545-1560): the 4-point function broadened to 8 meshwidths: first point NINT(t) - 4, the odd points
 * from r = (t - (lo + 3 + 1/2))/2, the even ones from r + 1/2, each scaled by 1/16; loop bounds clipped, weights indexed
 * from the unclipped first point, products w0*(w1*w2) (tensor style). */
static void stencil_ib_4_w8(double Xs, double x_lower, double dx, int ilower, int iupper, int g, stencil1d* s)
{
    const double X_o_dx = (Xs - x_lower) / dx;
    const int ic_lower = nint_f(X_o_dx) + ilower - 4;
    const int ic_upper = ic_lower + 7;
    double r = 0.5 * (X_o_dx - ((double)(ic_lower + 3 - ilower) + 0.5));
    double q = sqrt(1.0 + 4.0 * r * (1.0 - r));
    s->w[1] = 0.0625 * (3.0 - 2.0 * r - q);
    s->w[3] = 0.0625 * (3.0 - 2.0 * r + q);
    s->w[5] = 0.0625 * (1.0 + 2.0 * r + q);
    s->w[7] = 0.0625 * (1.0 + 2.0 * r - q);
    r = r + 0.5;
    q = sqrt(1.0 + 4.0 * r * (1.0 - r));
    s->w[0] = 0.0625 * (3.0 - 2.0 * r - q);
    s->w[2] = 0.0625 * (3.0 - 2.0 * r + q);
    s->w[4] = 0.0625 * (1.0 + 2.0 * r + q);
    s->w[6] = 0.0625 * (1.0 + 2.0 * r - q);
    s->wbase = ic_lower;
    s->lo = (ic_lower > ilower - g) ? ic_lower : ilower - g;
    s->hi = (ic_upper < iupper + g) ? ic_upper : iupper + g;
}

static void make_stencil(int kernel,
                         int on_axis,
                         double Xs,
                         double Xraw,
                         double x_lower,
                         double dx,
                         int ilower,
                         int iupper,
                         int g,
                         stencil1d* s)
{
    switch (kernel)
    {
    case K_PIECEWISE_LINEAR:
        stencil_piecewise_linear(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_IB_4:
        stencil_ib_4(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_IB_6:
        stencil_ib_6(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_BSPLINE_3:
        stencil_bspline_3(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_IB_3:
        stencil_delta(ib_3_delta, 1, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_BSPLINE_5:
        stencil_delta(bspline_5_delta, 2, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_BSPLINE_6:
        stencil_delta(bspline_6_delta, 3, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_PIECEWISE_CUBIC:
        stencil_delta(piecewise_cubic_delta, 2, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_IB_5:
        stencil_ib_5(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_PIECEWISE_CONSTANT:
        stencil_piecewise_constant(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    /* composite B-splines (3d.f.m4:3510-5460): the index rule of the wider of the two (centred 3 / 5 points, sided 4 / 6
     * points); the first digit names the function along `axis`, the second the function of the other dimensions */
    case K_COMPOSITE_BSPLINE_32:
        stencil_delta(on_axis ? bspline_3_delta : piecewise_linear_delta, 1, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_23:
        stencil_delta(on_axis ? piecewise_linear_delta : bspline_3_delta, 1, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_43:
        stencil_delta(on_axis ? bspline_4_delta : bspline_3_delta, 2, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_34:
        stencil_delta(on_axis ? bspline_3_delta : bspline_4_delta, 2, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_54:
        stencil_delta(on_axis ? bspline_5_delta : bspline_4_delta, 2, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_45:
        stencil_delta(on_axis ? bspline_4_delta : bspline_5_delta, 2, 0, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_65:
        stencil_delta(on_axis ? bspline_6_delta : bspline_5_delta, 3, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_COMPOSITE_BSPLINE_56:
        stencil_delta(on_axis ? bspline_5_delta : bspline_6_delta, 3, 1, Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_DISCONTINUOUS_LINEAR:
        stencil_discontinuous_linear(on_axis, Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    case K_IB_4_W8:
        stencil_ib_4_w8(Xs, x_lower, dx, ilower, iupper, g, s);
        break;
    default:
        stencil_bspline_4(Xs, Xraw, x_lower, dx, ilower, iupper, g, s);
        break;
    }
}

static inline int is_tensor_style(int kernel)
{
    return kernel == K_IB_4 || kernel == K_IB_6 || kernel == K_IB_4_W8;
}

/*
 * Generic interpolate: V(d,s) = sum_stencil w * u(...,d) for every listed marker.
 * ndim = 2 or 3; for ndim == 2 the third dimension is a single plane.
 */
void le_oracle_interp(int kernel,
                      int ndim,
                      const double* dx,
                      const double* x_lower,
                      int depth,
                      const int* ilower,
                      const int* iupper,
                      const int* nugc,
                      const double* u,
                      const int* indices,
                      const double* Xshift,
                      int nindices,
                      const double* X,
                      double* V)
{
    ptrdiff_t n[3] = { 1, 1, 1 };
    int iglo[3] = { 0, 0, 0 };
    for (int d = 0; d < ndim; ++d)
    {
        n[d] = (ptrdiff_t)(iupper[d] - ilower[d] + 1 + 2 * nugc[d]);
        iglo[d] = ilower[d] - nugc[d];
    }
    const ptrdiff_t comp_stride = n[0] * n[1] * n[2];
    const int axis = kernel >> 8; /* the Fortran routines' `axis` argument (0 unless the caller set it) */
    kernel &= 0xff;
    const int tensor = is_tensor_style(kernel);
    for (int l = 0; l < nindices; ++l)
    {
        const int s = indices[l];
        stencil1d st[3];
        st[2].lo = st[2].hi = st[2].wbase = 0;
        st[2].w[0] = 1.0;
        for (int d = 0; d < ndim; ++d)
        {
            const double Xraw = X[(ptrdiff_t)ndim * s + d];
            const double Xs = Xraw + Xshift[(ptrdiff_t)ndim * l + d];
            make_stencil(kernel, d == axis, Xs, Xraw, x_lower[d], dx[d], ilower[d], iupper[d], nugc[d], &st[d]);
        }
        for (int d = 0; d < depth; ++d)
        {
            const double* ud = u + comp_stride * d;
            double acc = 0.0;
            for (int ic2 = st[2].lo; ic2 <= st[2].hi; ++ic2)
            {
                const double w2 = st[2].w[ic2 - st[2].wbase];
                for (int ic1 = st[1].lo; ic1 <= st[1].hi; ++ic1)
                {
                    const double w1 = st[1].w[ic1 - st[1].wbase];
                    /* TENSOR: w(i0,i1,i2) = w0*(w1*w2) (3d.f.m4:1297-1305); 2D: w0*w1. */
                    const double wyz = (ndim == 3) ? w1 * w2 : w1;
                    const double* row = ud + ((ptrdiff_t)(ic2 - iglo[2]) * n[1] + (ic1 - iglo[1])) * n[0] - iglo[0];
                    for (int ic0 = st[0].lo; ic0 <= st[0].hi; ++ic0)
                    {
                        const double w0 = st[0].w[ic0 - st[0].wbase];
                        if (tensor)
                            acc = acc + (w0 * wyz) * row[ic0];
                        else if (ndim == 3)
                            acc = acc + w0 * w1 * w2 * row[ic0]; /* 3d.f.m4:30-34 */
                        else
                            acc = acc + w0 * w1 * row[ic0]; /* 2d.f.m4:29-32 */
                    }
                }
            }
            V[(ptrdiff_t)depth * s + d] = acc;
        }
    }
}

/*
 * Generic spread: u(...,d) += w * V(d,s) / prod(dx) for every listed marker,
 * markers processed in list order (serial, as the Fortran does).
 */
void le_oracle_spread(int kernel,
                      int ndim,
                      const double* dx,
                      const double* x_lower,
                      int depth,
                      const int* indices,
                      const double* Xshift,
                      int nindices,
                      const double* X,
                      const double* V,
                      const int* ilower,
                      const int* iupper,
                      const int* nugc,
                      double* u)
{
    ptrdiff_t n[3] = { 1, 1, 1 };
    int iglo[3] = { 0, 0, 0 };
    for (int d = 0; d < ndim; ++d)
    {
        n[d] = (ptrdiff_t)(iupper[d] - ilower[d] + 1 + 2 * nugc[d]);
        iglo[d] = ilower[d] - nugc[d];
    }
    const ptrdiff_t comp_stride = n[0] * n[1] * n[2];
    const int axis = kernel >> 8;
    kernel &= 0xff;
    const int tensor = is_tensor_style(kernel);
    const double dxprod = (ndim == 3) ? dx[0] * dx[1] * dx[2] : dx[0] * dx[1];
    for (int l = 0; l < nindices; ++l)
    {
        const int s = indices[l];
        stencil1d st[3];
        st[2].lo = st[2].hi = st[2].wbase = 0;
        st[2].w[0] = 1.0;
        for (int d = 0; d < ndim; ++d)
        {
            const double Xraw = X[(ptrdiff_t)ndim * s + d];
            const double Xs = Xraw + Xshift[(ptrdiff_t)ndim * l + d];
            make_stencil(kernel, d == axis, Xs, Xraw, x_lower[d], dx[d], ilower[d], iupper[d], nugc[d], &st[d]);
        }
        for (int d = 0; d < depth; ++d)
        {
            double* ud = u + comp_stride * d;
            const double Vds = V[(ptrdiff_t)depth * s + d];
            for (int ic2 = st[2].lo; ic2 <= st[2].hi; ++ic2)
            {
                const double w2 = st[2].w[ic2 - st[2].wbase];
                /* TENSOR 3D: wz = w2/(dx0*dx1*dx2) (3d.f.m4:1439) */
                const double wz = w2 / dxprod;
                for (int ic1 = st[1].lo; ic1 <= st[1].hi; ++ic1)
                {
                    const double w1 = st[1].w[ic1 - st[1].wbase];
                    /* TENSOR 2D: wy = w1/(dx0*dx1) (2d.f.m4:1356) */
                    const double wyz = (ndim == 3) ? w1 * wz : w1 / dxprod;
                    double* row = ud + ((ptrdiff_t)(ic2 - iglo[2]) * n[1] + (ic1 - iglo[1])) * n[0] - iglo[0];
                    for (int ic0 = st[0].lo; ic0 <= st[0].hi; ++ic0)
                    {
                        const double w0 = st[0].w[ic0 - st[0].wbase];
                        if (tensor)
                            row[ic0] = row[ic0] + (w0 * wyz) * Vds;
                        else if (ndim == 3)
                            row[ic0] = row[ic0] + (w0 * w1 * w2 * Vds / dxprod); /* 3d.f.m4:61-65 */
                        else
                            row[ic0] = row[ic0] + (w0 * w1 * Vds / dxprod); /* 2d.f.m4:54-57 */
                    }
                }
            }
        }
    }
}

/*
 * Fortran-ABI entry points (seam B4): the exact symbol names and argument
 * orders the reference declares at LEInteractor.cpp:237-1518 (mangling:
 * lower case + trailing underscore, CMakeLists.txt:77-83), so this file could
 * be link-substituted into a real IBAMR build for cross-checking.
 */
#define DEFINE_3D(NAME, KERNEL)                                                                                        \
    void lagrangian_##NAME##_interp3d_(const double* dx,                                                               \
                                       const double* x_lower,                                                          \
                                       const double* x_upper,                                                          \
                                       const int* depth,                                                               \
                                       const int* ilower0,                                                             \
                                       const int* iupper0,                                                             \
                                       const int* ilower1,                                                             \
                                       const int* iupper1,                                                             \
                                       const int* ilower2,                                                             \
                                       const int* iupper2,                                                             \
                                       const int* nugc0,                                                               \
                                       const int* nugc1,                                                               \
                                       const int* nugc2,                                                               \
                                       const double* u,                                                                \
                                       const int* indices,                                                             \
                                       const double* Xshift,                                                           \
                                       const int* nindices,                                                            \
                                       const double* X,                                                                \
                                       double* V)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[3] = { *ilower0, *ilower1, *ilower2 }, iu[3] = { *iupper0, *iupper1, *iupper2 };                  \
        const int ng[3] = { *nugc0, *nugc1, *nugc2 };                                                                  \
        le_oracle_interp(KERNEL, 3, dx, x_lower, *depth, il, iu, ng, u, indices, Xshift, *nindices, X, V);             \
    }                                                                                                                  \
    void lagrangian_##NAME##_spread3d_(const double* dx,                                                               \
                                       const double* x_lower,                                                          \
                                       const double* x_upper,                                                          \
                                       const int* depth,                                                               \
                                       const int* indices,                                                             \
                                       const double* Xshift,                                                           \
                                       const int* nindices,                                                            \
                                       const double* X,                                                                \
                                       const double* V,                                                                \
                                       const int* ilower0,                                                             \
                                       const int* iupper0,                                                             \
                                       const int* ilower1,                                                             \
                                       const int* iupper1,                                                             \
                                       const int* ilower2,                                                             \
                                       const int* iupper2,                                                             \
                                       const int* nugc0,                                                               \
                                       const int* nugc1,                                                               \
                                       const int* nugc2,                                                               \
                                       double* u)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[3] = { *ilower0, *ilower1, *ilower2 }, iu[3] = { *iupper0, *iupper1, *iupper2 };                  \
        const int ng[3] = { *nugc0, *nugc1, *nugc2 };                                                                  \
        le_oracle_spread(KERNEL, 3, dx, x_lower, *depth, indices, Xshift, *nindices, X, V, il, iu, ng, u);             \
    }

#define DEFINE_2D(NAME, KERNEL)                                                                                        \
    void lagrangian_##NAME##_interp2d_(const double* dx,                                                               \
                                       const double* x_lower,                                                          \
                                       const double* x_upper,                                                          \
                                       const int* depth,                                                               \
                                       const int* ilower0,                                                             \
                                       const int* iupper0,                                                             \
                                       const int* ilower1,                                                             \
                                       const int* iupper1,                                                             \
                                       const int* nugc0,                                                               \
                                       const int* nugc1,                                                               \
                                       const double* u,                                                                \
                                       const int* indices,                                                             \
                                       const double* Xshift,                                                           \
                                       const int* nindices,                                                            \
                                       const double* X,                                                                \
                                       double* V)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[2] = { *ilower0, *ilower1 }, iu[2] = { *iupper0, *iupper1 };                                      \
        const int ng[2] = { *nugc0, *nugc1 };                                                                          \
        le_oracle_interp(KERNEL, 2, dx, x_lower, *depth, il, iu, ng, u, indices, Xshift, *nindices, X, V);             \
    }                                                                                                                  \
    void lagrangian_##NAME##_spread2d_(const double* dx,                                                               \
                                       const double* x_lower,                                                          \
                                       const double* x_upper,                                                          \
                                       const int* depth,                                                               \
                                       const int* indices,                                                             \
                                       const double* Xshift,                                                           \
                                       const int* nindices,                                                            \
                                       const double* X,                                                                \
                                       const double* V,                                                                \
                                       const int* ilower0,                                                             \
                                       const int* iupper0,                                                             \
                                       const int* ilower1,                                                             \
                                       const int* iupper1,                                                             \
                                       const int* nugc0,                                                               \
                                       const int* nugc1,                                                               \
                                       double* u)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[2] = { *ilower0, *ilower1 }, iu[2] = { *iupper0, *iupper1 };                                      \
        const int ng[2] = { *nugc0, *nugc1 };                                                                          \
        le_oracle_spread(KERNEL, 2, dx, x_lower, *depth, indices, Xshift, *nindices, X, V, il, iu, ng, u);             \
    }

DEFINE_3D(piecewise_linear, K_PIECEWISE_LINEAR)
DEFINE_3D(ib_4, K_IB_4)
DEFINE_3D(ib_6, K_IB_6)
DEFINE_3D(bspline_3, K_BSPLINE_3)
DEFINE_3D(bspline_4, K_BSPLINE_4)
DEFINE_3D(ib_3, K_IB_3)
DEFINE_3D(bspline_5, K_BSPLINE_5)
DEFINE_3D(bspline_6, K_BSPLINE_6)
DEFINE_3D(piecewise_cubic, K_PIECEWISE_CUBIC)
DEFINE_3D(ib_5, K_IB_5)
DEFINE_3D(piecewise_constant, K_PIECEWISE_CONSTANT)
DEFINE_2D(piecewise_linear, K_PIECEWISE_LINEAR)
DEFINE_2D(ib_4, K_IB_4)
DEFINE_2D(ib_6, K_IB_6)
DEFINE_2D(bspline_3, K_BSPLINE_3)
DEFINE_2D(bspline_4, K_BSPLINE_4)
DEFINE_2D(ib_3, K_IB_3)
DEFINE_2D(bspline_5, K_BSPLINE_5)
DEFINE_2D(bspline_6, K_BSPLINE_6)
DEFINE_2D(piecewise_cubic, K_PIECEWISE_CUBIC)
DEFINE_2D(ib_5, K_IB_5)
DEFINE_2D(piecewise_constant, K_PIECEWISE_CONSTANT)

/* The axis-dependent routines take `axis` right after `depth` (3d.f.m4:237-243, 3510-3516; 2d twins). */
#define DEFINE_AXIS_3D(NAME, KERNEL)                                                                                   \
    void lagrangian_##NAME##_interp3d_(const double* dx, const double* x_lower, const double* x_upper, const int* depth, \
                                       const int* axis, const int* ilower0, const int* iupper0, const int* ilower1,    \
                                       const int* iupper1, const int* ilower2, const int* iupper2, const int* nugc0,   \
                                       const int* nugc1, const int* nugc2, const double* u, const int* indices,        \
                                       const double* Xshift, const int* nindices, const double* X, double* V)          \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[3] = { *ilower0, *ilower1, *ilower2 }, iu[3] = { *iupper0, *iupper1, *iupper2 };                  \
        const int ng[3] = { *nugc0, *nugc1, *nugc2 };                                                                  \
        le_oracle_interp(KERNEL | (*axis << 8), 3, dx, x_lower, *depth, il, iu, ng, u, indices, Xshift, *nindices, X, V); \
    }                                                                                                                  \
    void lagrangian_##NAME##_spread3d_(const double* dx, const double* x_lower, const double* x_upper, const int* depth, \
                                       const int* axis, const int* indices, const double* Xshift, const int* nindices, \
                                       const double* X, const double* V, const int* ilower0, const int* iupper0,       \
                                       const int* ilower1, const int* iupper1, const int* ilower2, const int* iupper2, \
                                       const int* nugc0, const int* nugc1, const int* nugc2, double* u)                \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[3] = { *ilower0, *ilower1, *ilower2 }, iu[3] = { *iupper0, *iupper1, *iupper2 };                  \
        const int ng[3] = { *nugc0, *nugc1, *nugc2 };                                                                  \
        le_oracle_spread(KERNEL | (*axis << 8), 3, dx, x_lower, *depth, indices, Xshift, *nindices, X, V, il, iu, ng, u); \
    }
#define DEFINE_AXIS_2D(NAME, KERNEL)                                                                                   \
    void lagrangian_##NAME##_interp2d_(const double* dx, const double* x_lower, const double* x_upper, const int* depth, \
                                       const int* axis, const int* ilower0, const int* iupper0, const int* ilower1,    \
                                       const int* iupper1, const int* nugc0, const int* nugc1, const double* u,        \
                                       const int* indices, const double* Xshift, const int* nindices, const double* X, \
                                       double* V)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[2] = { *ilower0, *ilower1 }, iu[2] = { *iupper0, *iupper1 };                                      \
        const int ng[2] = { *nugc0, *nugc1 };                                                                          \
        le_oracle_interp(KERNEL | (*axis << 8), 2, dx, x_lower, *depth, il, iu, ng, u, indices, Xshift, *nindices, X, V); \
    }                                                                                                                  \
    void lagrangian_##NAME##_spread2d_(const double* dx, const double* x_lower, const double* x_upper, const int* depth, \
                                       const int* axis, const int* indices, const double* Xshift, const int* nindices, \
                                       const double* X, const double* V, const int* ilower0, const int* iupper0,       \
                                       const int* ilower1, const int* iupper1, const int* nugc0, const int* nugc1,     \
                                       double* u)                                                                      \
    {                                                                                                                  \
        (void)x_upper;                                                                                                 \
        const int il[2] = { *ilower0, *ilower1 }, iu[2] = { *iupper0, *iupper1 };                                      \
        const int ng[2] = { *nugc0, *nugc1 };                                                                          \
        le_oracle_spread(KERNEL | (*axis << 8), 2, dx, x_lower, *depth, indices, Xshift, *nindices, X, V, il, iu, ng, u); \
    }
DEFINE_3D(ib_4_w8, K_IB_4_W8)
DEFINE_2D(ib_4_w8, K_IB_4_W8)
#define DEFINE_AXIS(NAME, KERNEL) DEFINE_AXIS_3D(NAME, KERNEL) DEFINE_AXIS_2D(NAME, KERNEL)
DEFINE_AXIS(composite_bspline_32, K_COMPOSITE_BSPLINE_32)
DEFINE_AXIS(composite_bspline_23, K_COMPOSITE_BSPLINE_23)
DEFINE_AXIS(composite_bspline_43, K_COMPOSITE_BSPLINE_43)
DEFINE_AXIS(composite_bspline_34, K_COMPOSITE_BSPLINE_34)
DEFINE_AXIS(composite_bspline_54, K_COMPOSITE_BSPLINE_54)
DEFINE_AXIS(composite_bspline_45, K_COMPOSITE_BSPLINE_45)
DEFINE_AXIS(composite_bspline_65, K_COMPOSITE_BSPLINE_65)
DEFINE_AXIS(composite_bspline_56, K_COMPOSITE_BSPLINE_56)
DEFINE_AXIS(discontinuous_linear, K_DISCONTINUOUS_LINEAR)
