/*
 * oracle/le_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the host-side logic that wraps the Fortran kernels in the
 * reference (parity status: PINNED through tests/test_oracle_golden.py, see
 * le_kernels.c header):
 *
 *   le_oracle_get_cell_index      IndexUtilities::getCellIndex
 *                                 ibtk/include/ibtk/private/IndexUtilities-inl.h:50-75
 *   le_oracle_indices_in_box      LEInteractor::buildLocalIndices (position form)
 *                                 ibtk/src/lagrangian/LEInteractor.cpp:6088-6126
 *   le_oracle_side_interp/spread  per-axis SideData decomposition
 *                                 LEInteractor.cpp:2402-2489 (interp), :3627-3714 (spread),
 *                                 position-only forms :3045-3127, :4188-4263
 *   le_oracle_patch_lists         LIndexSetData::cacheLocalIndices (which markers a patch
 *                                 sees, interior vs ghost, periodic shifts)
 *                                 ibtk/src/lagrangian/LIndexSetData.cpp:53-141, with the
 *                                 canonical order (cell k-j-i, then Lagrangian index;
 *                                 LDataManager.cpp:1505, 2897-2911)
 *   le_oracle_wrap_positions      pre-binning wrap/clamp, LDataManager.cpp:1397-1421
 *   le_oracle_baseline_*          the reference's parallel model (one worker per patch,
 *                                 private arrays, redundant ghost-region spreading,
 *                                 LDataManager.cpp:623-652 / :748-802) on OpenMP threads,
 *                                 used ONLY as bench.py's cpu_baseline / --impl reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

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
                      double* V);
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
                      double* u);

/* IndexUtilities-inl.h:62-73, vectorised over n points (X is AoS [n][ndim]). */
void le_oracle_get_cell_index(int ndim,
                              int n,
                              const double* X,
                              const double* x_lower,
                              const double* x_upper,
                              const double* dx,
                              const int* ilower,
                              const int* iupper,
                              int* idx)
{
    for (int k = 0; k < n; ++k)
    {
        for (int d = 0; d < ndim; ++d)
        {
            const double dX_lower = X[(size_t)ndim * k + d] - x_lower[d];
            const double dX_upper = X[(size_t)ndim * k + d] - x_upper[d];
            if (fabs(dX_lower) <= fabs(dX_upper))
                idx[(size_t)ndim * k + d] = ilower[d] + (int)floor(dX_lower / dx[d]);
            else
                idx[(size_t)ndim * k + d] = iupper[d] + (int)floor(dX_upper / dx[d]) + 1;
        }
    }
}

/* LEInteractor.cpp:6110-6124: indices (in input order) of the points whose cell lies in box. */
int le_oracle_indices_in_box(int ndim,
                             int n,
                             const double* X,
                             const double* x_lower,
                             const double* x_upper,
                             const double* dx,
                             const int* patch_lower,
                             const int* patch_upper,
                             const int* box_lower,
                             const int* box_upper,
                             int* out_indices)
{
    int count = 0;
    for (int k = 0; k < n; ++k)
    {
        int c[3];
        le_oracle_get_cell_index(ndim, 1, X + (size_t)ndim * k, x_lower, x_upper, dx, patch_lower, patch_upper, c);
        int inside = 1;
        for (int d = 0; d < ndim; ++d) inside = inside && c[d] >= box_lower[d] && c[d] <= box_upper[d];
        if (inside) out_indices[count++] = k;
    }
    return count;
}

/* LDataManager.cpp:1397-1421.  Returns the number of points that left a non-periodic domain
 * (the reference aborts on those when error_if_points_leave_domain is set). */
int le_oracle_wrap_positions(int ndim,
                             int n,
                             double* X,
                             const double* domain_x_lower,
                             const double* domain_x_upper,
                             const int* periodic)
{
    const double TOL = sqrt(2.220446049250313e-16); /* LDataManager.cpp:150 */
    int escaped = 0;
    for (int k = 0; k < n; ++k)
    {
        for (int d = 0; d < ndim; ++d)
        {
            double* x = &X[(size_t)ndim * k + d];
            if (periodic[d])
            {
                const double L = domain_x_upper[d] - domain_x_lower[d];
                while (*x < domain_x_lower[d]) *x += L;
                while (*x >= domain_x_upper[d]) *x -= L;
            }
            else
            {
                if (*x < domain_x_lower[d] || *x > domain_x_upper[d]) ++escaped;
                *x = fmax(*x, domain_x_lower[d]);
                *x = fmin(*x, domain_x_upper[d] - (domain_x_upper[d] - domain_x_lower[d]) * TOL);
            }
        }
    }
    return escaped;
}

/* Side box of axis `axis` for a cell box (SideGeometry::toSideBox): upper[axis] + 1. */
static void side_box(int ndim, int axis, const int* lo, const int* hi, int* slo, int* shi)
{
    for (int d = 0; d < ndim; ++d)
    {
        slo[d] = lo[d];
        shi[d] = hi[d] + (d == axis ? 1 : 0);
    }
}

/*
 * LEInteractor.cpp:2454-2486 (index-set form; the position-only form :3094-3124 is the same
 * loop with zero shifts).  u[axis] is the Fortran array of the axis-normal component over the
 * side box grown by gcw.  Q is AoS [*][ndim]; only listed markers are written.
 */
void le_oracle_side_interp(int kernel,
                           int ndim,
                           const double* x_lower,
                           const double* dx,
                           const int* patch_lower,
                           const int* patch_upper,
                           const int* gcw,
                           const double* const* u,
                           const int* indices,
                           const double* shifts,
                           int nindices,
                           const double* X,
                           double* Q)
{
    if (nindices == 0) return;
    int local_sz = 0;
    for (int l = 0; l < nindices; ++l)
        if (indices[l] + 1 > local_sz) local_sz = indices[l] + 1;
    double* Q_axis = (double*)malloc(sizeof(double) * (size_t)local_sz);
    for (int axis = 0; axis < ndim; ++axis)
    {
        double x_lower_axis[3];
        int slo[3], shi[3];
        for (int d = 0; d < ndim; ++d) x_lower_axis[d] = x_lower[d];
        x_lower_axis[axis] -= 0.5 * dx[axis]; /* LEInteractor.cpp:2464 */
        side_box(ndim, axis, patch_lower, patch_upper, slo, shi);
        le_oracle_interp(kernel | (axis << 8), ndim, dx, x_lower_axis, 1, slo, shi, gcw, u[axis], indices, shifts, nindices, X, Q_axis);
        for (int l = 0; l < nindices; ++l) Q[(size_t)ndim * indices[l] + axis] = Q_axis[indices[l]];
    }
    free(Q_axis);
}

/* LEInteractor.cpp:3676-3711 (and :4229-4260). */
void le_oracle_side_spread(int kernel,
                           int ndim,
                           const double* x_lower,
                           const double* dx,
                           const int* patch_lower,
                           const int* patch_upper,
                           const int* gcw,
                           double* const* u,
                           const int* indices,
                           const double* shifts,
                           int nindices,
                           const double* X,
                           const double* Q)
{
    if (nindices == 0) return;
    int local_sz = 0;
    for (int l = 0; l < nindices; ++l)
        if (indices[l] + 1 > local_sz) local_sz = indices[l] + 1;
    double* Q_axis = (double*)malloc(sizeof(double) * (size_t)local_sz);
    for (int axis = 0; axis < ndim; ++axis)
    {
        double x_lower_axis[3];
        int slo[3], shi[3];
        for (int d = 0; d < ndim; ++d) x_lower_axis[d] = x_lower[d];
        x_lower_axis[axis] -= 0.5 * dx[axis]; /* LEInteractor.cpp:3689 */
        side_box(ndim, axis, patch_lower, patch_upper, slo, shi);
        for (int l = 0; l < nindices; ++l) Q_axis[indices[l]] = Q[(size_t)ndim * indices[l] + axis];
        le_oracle_spread(kernel | (axis << 8), ndim, dx, x_lower_axis, 1, indices, shifts, nindices, X, Q_axis, slo, shi, gcw, u[axis]);
    }
    free(Q_axis);
}

/*
 * Which markers does a patch see?  (LIndexSetData.cpp:76-137 restated on flat arrays.)
 * cells[n][ndim] are level cell indices (from le_oracle_get_cell_index on the level
 * geometry).  A marker is listed once per periodic image whose cell lies in the patch box
 * grown by gcw; `interior` is 1 iff the (unshifted) cell lies in the patch box itself.
 * The periodic offset of an image is expressed exactly as the reference does: the shift that
 * must be ADDED to X to see the marker from this patch, offset*dx (:89-101).
 * Output is in canonical order: cell (k, j, i) of the image with i fastest, then marker index.
 * Returns the number of entries (call with out_idx == NULL to count).
 */
typedef struct
{
    long long key;
    int idx;
    int off[3];
    int interior;
} patch_entry;

static int cmp_entry(const void* a, const void* b)
{
    const patch_entry* x = (const patch_entry*)a;
    const patch_entry* y = (const patch_entry*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

int le_oracle_patch_lists(int ndim,
                          int n,
                          const int* cells,
                          const int* patch_lower,
                          const int* patch_upper,
                          const int* gcw,
                          const int* domain_lower,
                          const int* domain_ncells,
                          const int* periodic,
                          const double* dx,
                          int* out_idx,
                          double* out_shift,
                          int* out_interior)
{
    int glo[3] = { 0, 0, 0 }, ghi[3] = { 0, 0, 0 }, gn[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d)
    {
        glo[d] = patch_lower[d] - gcw[d];
        ghi[d] = patch_upper[d] + gcw[d];
        gn[d] = ghi[d] - glo[d] + 1;
    }
    (void)domain_lower;
    size_t cap = 1024, cnt = 0;
    patch_entry* e = (patch_entry*)malloc(cap * sizeof(patch_entry));
    for (int k = 0; k < n; ++k)
    {
        int omin[3] = { 0, 0, 0 }, omax[3] = { 0, 0, 0 };
        for (int d = 0; d < ndim; ++d)
            if (periodic[d])
            {
                omin[d] = -1;
                omax[d] = 1;
            }
        for (int o2 = omin[2]; o2 <= omax[2]; ++o2)
            for (int o1 = omin[1]; o1 <= omax[1]; ++o1)
                for (int o0 = omin[0]; o0 <= omax[0]; ++o0)
                {
                    const int o[3] = { o0, o1, o2 };
                    int c[3] = { 0, 0, 0 };
                    int in_ghost = 1, in_patch = 1;
                    for (int d = 0; d < ndim; ++d)
                    {
                        c[d] = cells[(size_t)ndim * k + d] + o[d] * domain_ncells[d];
                        in_ghost = in_ghost && c[d] >= glo[d] && c[d] <= ghi[d];
                        in_patch = in_patch && c[d] >= patch_lower[d] && c[d] <= patch_upper[d];
                    }
                    if (!in_ghost) continue;
                    if (cnt == cap)
                    {
                        cap *= 2;
                        e = (patch_entry*)realloc(e, cap * sizeof(patch_entry));
                    }
                    patch_entry* p = &e[cnt++];
                    p->key = ((long long)(c[2] - glo[2]) * gn[1] + (c[1] - glo[1])) * gn[0] + (c[0] - glo[0]);
                    p->idx = k;
                    p->interior = in_patch;
                    for (int d = 0; d < 3; ++d) p->off[d] = (d < ndim) ? o[d] * domain_ncells[d] : 0;
                }
    }
    qsort(e, cnt, sizeof(patch_entry), cmp_entry);
    if (out_idx)
    {
        for (size_t i = 0; i < cnt; ++i)
        {
            out_idx[i] = e[i].idx;
            out_interior[i] = e[i].interior;
            for (int d = 0; d < ndim; ++d) out_shift[(size_t)ndim * i + d] = (double)e[i].off[d] * dx[d];
        }
    }
    free(e);
    return (int)cnt;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline: the reference's parallel model on OpenMP threads.
 *
 * A periodic N[0] x N[1] x N[2] level is cut into np[0] x np[1] x np[2] equal patches (the
 * SAMRAI box decomposition an MPI run would use, one patch per "rank").  Each worker owns
 * private side-centred arrays for its patch with gcw ghost layers, spreads from every marker
 * in its GHOST box (redundant spreading, LDataManager.cpp:623-652) and interpolates to the
 * markers in its INTERIOR box (:748-802).  Ghost cells of u are filled analytically by the
 * caller through the `u_fill` sampling below (the fill itself is SAMRAI's job and is untimed,
 * exactly as the GPU arm starts from filled ghosts).
 * ------------------------------------------------------------------------------------------ */
typedef struct
{
    int ndim;
    int npatch;
    int gcw[3];
    int* plo;       /* [npatch][3] */
    int* phi;       /* [npatch][3] */
    double* xlo;    /* [npatch][3] */
    double dx[3];
    double** u;     /* [npatch*ndim] */
    double** f;     /* [npatch*ndim] */
    size_t* usize;  /* [npatch*ndim] */
    int* nall;      /* [npatch] entries in ghost-box list */
    int** all_idx;  /* [npatch] */
    double** all_shift;
    int* nint;      /* interior list */
    int** int_idx;
    double** int_shift;
} baseline_t;

static double splitmix_unit(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x = x ^ (x >> 31);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

/* u_a(x) = sin(2 pi x_a) cos(2 pi x_{a+1}) + 1e-3 * noise(global side index), SURVEY 8(d). */
static double u_field(int ndim, int axis, const double* x, const long long gidx, unsigned long long seed)
{
    const double twopi = 6.283185307179586476925286766559;
    const int b = (axis + 1) % ndim;
    return sin(twopi * x[axis]) * cos(twopi * x[b]) + 1e-3 * splitmix_unit(seed ^ (unsigned long long)gidx);
}

void le_oracle_baseline_destroy(baseline_t* b);

baseline_t* le_oracle_baseline_create(int ndim,
                                      const int* N,
                                      const int* np,
                                      int gcw,
                                      const double* domain_x_lower,
                                      const double* domain_x_upper,
                                      int nmarkers,
                                      const double* X,
                                      unsigned long long field_seed)
{
    baseline_t* b = (baseline_t*)calloc(1, sizeof(baseline_t));
    b->ndim = ndim;
    int npz = ndim == 3 ? np[2] : 1;
    b->npatch = np[0] * np[1] * npz;
    int Nn[3] = { N[0], N[1], ndim == 3 ? N[2] : 1 };
    int dom_lo[3] = { 0, 0, 0 }, dom_hi[3] = { 0, 0, 0 }, periodic[3] = { 1, 1, 1 };
    for (int d = 0; d < 3; ++d)
    {
        b->gcw[d] = d < ndim ? gcw : 0;
        dom_hi[d] = Nn[d] - 1;
    }
    for (int d = 0; d < ndim; ++d) b->dx[d] = (domain_x_upper[d] - domain_x_lower[d]) / (double)N[d];
    b->plo = (int*)calloc((size_t)b->npatch * 3, sizeof(int));
    b->phi = (int*)calloc((size_t)b->npatch * 3, sizeof(int));
    b->xlo = (double*)calloc((size_t)b->npatch * 3, sizeof(double));
    b->u = (double**)calloc((size_t)b->npatch * ndim, sizeof(double*));
    b->f = (double**)calloc((size_t)b->npatch * ndim, sizeof(double*));
    b->usize = (size_t*)calloc((size_t)b->npatch * ndim, sizeof(size_t));
    b->nall = (int*)calloc(b->npatch, sizeof(int));
    b->nint = (int*)calloc(b->npatch, sizeof(int));
    b->all_idx = (int**)calloc(b->npatch, sizeof(int*));
    b->int_idx = (int**)calloc(b->npatch, sizeof(int*));
    b->all_shift = (double**)calloc(b->npatch, sizeof(double*));
    b->int_shift = (double**)calloc(b->npatch, sizeof(double*));

    int* cells = (int*)malloc(sizeof(int) * (size_t)ndim * nmarkers);
    le_oracle_get_cell_index(ndim, nmarkers, X, domain_x_lower, domain_x_upper, b->dx, dom_lo, dom_hi, cells);

    int p = 0;
    for (int pz = 0; pz < npz; ++pz)
        for (int py = 0; py < np[1]; ++py)
            for (int px = 0; px < np[0]; ++px, ++p)
            {
                const int pc[3] = { px, py, pz };
                for (int d = 0; d < ndim; ++d)
                {
                    const int w = Nn[d] / np[d];
                    b->plo[3 * p + d] = pc[d] * w;
                    b->phi[3 * p + d] = (pc[d] + 1) * w - 1;
                    b->xlo[3 * p + d] = domain_x_lower[d] + b->dx[d] * (double)b->plo[3 * p + d];
                }
            }
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < b->npatch; ++q)
    {
        const int* plo = &b->plo[3 * q];
        const int* phi = &b->phi[3 * q];
        /* marker lists (LIndexSetData::cacheLocalIndices) */
        const int cnt =
            le_oracle_patch_lists(ndim, nmarkers, cells, plo, phi, b->gcw, dom_lo, Nn, periodic, b->dx, NULL, NULL, NULL);
        int* idx = (int*)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
        int* interior = (int*)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
        double* shift = (double*)malloc(sizeof(double) * (size_t)ndim * (size_t)(cnt > 0 ? cnt : 1));
        le_oracle_patch_lists(ndim, nmarkers, cells, plo, phi, b->gcw, dom_lo, Nn, periodic, b->dx, idx, shift, interior);
        b->nall[q] = cnt;
        b->all_idx[q] = idx;
        b->all_shift[q] = shift;
        int ni = 0;
        for (int i = 0; i < cnt; ++i) ni += interior[i];
        b->nint[q] = ni;
        b->int_idx[q] = (int*)malloc(sizeof(int) * (size_t)(ni > 0 ? ni : 1));
        b->int_shift[q] = (double*)malloc(sizeof(double) * (size_t)ndim * (size_t)(ni > 0 ? ni : 1));
        ni = 0;
        for (int i = 0; i < cnt; ++i)
            if (interior[i])
            {
                b->int_idx[q][ni] = idx[i];
                for (int d = 0; d < ndim; ++d) b->int_shift[q][(size_t)ndim * ni + d] = shift[(size_t)ndim * i + d];
                ++ni;
            }
        free(interior);
        /* arrays */
        for (int axis = 0; axis < ndim; ++axis)
        {
            ptrdiff_t n[3] = { 1, 1, 1 };
            for (int d = 0; d < ndim; ++d) n[d] = phi[d] - plo[d] + 1 + (d == axis ? 1 : 0) + 2 * b->gcw[d];
            const size_t sz = (size_t)(n[0] * n[1] * n[2]);
            b->usize[q * ndim + axis] = sz;
            double* u = (double*)malloc(sizeof(double) * sz);
            double* f = (double*)calloc(sz, sizeof(double));
            for (ptrdiff_t k = 0; k < n[2]; ++k)
                for (ptrdiff_t j = 0; j < n[1]; ++j)
                    for (ptrdiff_t i = 0; i < n[0]; ++i)
                    {
                        const ptrdiff_t ii[3] = { i, j, k };
                        double x[3] = { 0, 0, 0 };
                        long long g = 0, mul = 1;
                        for (int d = 0; d < ndim; ++d)
                        {
                            const long long gi = (long long)plo[d] - b->gcw[d] + ii[d];
                            const long long gw = ((gi % Nn[d]) + Nn[d]) % Nn[d]; /* periodic image */
                            x[d] = domain_x_lower[d] + b->dx[d] * ((double)gw + (d == axis ? 0.0 : 0.5));
                            g += gw * mul;
                            mul *= Nn[d];
                        }
                        u[(k * n[1] + j) * n[0] + i] = u_field(ndim, axis, x, g, field_seed + 11ULL + (unsigned)axis);
                    }
            b->u[q * ndim + axis] = u;
            b->f[q * ndim + axis] = f;
        }
    }
    free(cells);
    return b;
}

/* One pass of the hot path: spread F (ghost-box markers) then interpolate U (interior markers). */
void le_oracle_baseline_step(baseline_t* b, int kernel, const double* X, const double* F, double* U)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < b->npatch; ++q)
    {
        le_oracle_side_spread(kernel,
                              b->ndim,
                              &b->xlo[3 * q],
                              b->dx,
                              &b->plo[3 * q],
                              &b->phi[3 * q],
                              b->gcw,
                              &b->f[q * b->ndim],
                              b->all_idx[q],
                              b->all_shift[q],
                              b->nall[q],
                              X,
                              F);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < b->npatch; ++q)
    {
        le_oracle_side_interp(kernel,
                              b->ndim,
                              &b->xlo[3 * q],
                              b->dx,
                              &b->plo[3 * q],
                              &b->phi[3 * q],
                              b->gcw,
                              (const double* const*)&b->u[q * b->ndim],
                              b->int_idx[q],
                              b->int_shift[q],
                              b->nint[q],
                              X,
                              U);
    }
}

void le_oracle_baseline_zero_f(baseline_t* b)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < b->npatch * b->ndim; ++q) memset(b->f[q], 0, sizeof(double) * b->usize[q]);
}

/* Copy of a patch's arrays for checking (axis-normal component `axis` of patch q). */
size_t le_oracle_baseline_array_size(baseline_t* b, int q, int axis)
{
    return b->usize[q * b->ndim + axis];
}
const double* le_oracle_baseline_f(baseline_t* b, int q, int axis)
{
    return b->f[q * b->ndim + axis];
}
const double* le_oracle_baseline_u(baseline_t* b, int q, int axis)
{
    return b->u[q * b->ndim + axis];
}
int le_oracle_baseline_npatch(baseline_t* b)
{
    return b->npatch;
}
/* torchrun exports OMP_NUM_THREADS=1; the baseline must use the cores the process may run on. */
void le_oracle_baseline_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int le_oracle_baseline_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void le_oracle_baseline_destroy(baseline_t* b)
{
    if (!b) return;
    for (int q = 0; q < b->npatch; ++q)
    {
        free(b->all_idx[q]);
        free(b->all_shift[q]);
        free(b->int_idx[q]);
        free(b->int_shift[q]);
    }
    for (int q = 0; q < b->npatch * b->ndim; ++q)
    {
        free(b->u[q]);
        free(b->f[q]);
    }
    free(b->plo);
    free(b->phi);
    free(b->xlo);
    free(b->u);
    free(b->f);
    free(b->usize);
    free(b->nall);
    free(b->nint);
    free(b->all_idx);
    free(b->int_idx);
    free(b->all_shift);
    free(b->int_shift);
    free(b);
}
