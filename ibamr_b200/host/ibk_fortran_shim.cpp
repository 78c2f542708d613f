// ibk_fortran_shim.cpp -- link-substitution shim for seam B4: defines the Fortran symbols
// LEInteractor.cpp declares (LEInteractor.cpp:237-1518, mangling lower case + '_', CMakeLists.txt:77-83)
// for the five in-scope kernels and forwards them to libibk.so.  Linking this object in place of
// lagrangian_interaction{2,3}d.f.m4's objects moves the arithmetic of an UNMODIFIED LEInteractor to the
// GPU (one call per SideData axis, host arrays in, host arrays out).  Argument orders:
// lagrangian_interaction3d.f.m4:1203-1209 (interp), :1344-1350 (spread).
#include <cstdio>
#include <cstdlib>

#include "../../include/ibk.h"

namespace
{
ibk_ctx* ctx()
{
    static ibk_ctx* c = nullptr;
    if (!c && ibk_ctx_create(0, &c) != IBK_OK)
    {
        std::fprintf(stderr, "ibk_fortran_shim: no CUDA device (libibk has no CPU fallback)\n");
        std::abort(); // TBOX_ERROR -> MPI_Abort in the reference
    }
    return c;
}
void check(int rc)
{
    if (rc != IBK_OK)
    {
        std::fprintf(stderr, "ibk_fortran_shim: %s\n", ibk_last_error(ctx()));
        std::abort();
    }
}
int max_index(const int* indices, int n)
{
    int m = -1;
    for (int i = 0; i < n; ++i) m = indices[i] > m ? indices[i] : m;
    return m + 1;
}
ibk_array_desc desc(int ndim, const double* dx, const double* x_lower, const double* x_upper, int depth, const int* il,
                    const int* iu, const int* ng)
{
    ibk_array_desc d{};
    d.ndim = ndim;
    d.depth = depth;
    for (int k = 0; k < ndim; ++k)
    {
        d.dx[k] = dx[k];
        d.x_lower[k] = x_lower[k];
        d.x_upper[k] = x_upper[k];
        d.ilower[k] = il[k];
        d.iupper[k] = iu[k];
        d.nugc[k] = ng[k];
    }
    return d;
}
} // namespace

#define IBK_SHIM_3D(NAME, KERNEL)                                                                                      \
    extern "C" void lagrangian_##NAME##_interp3d_(                                                                     \
        const double* dx, const double* x_lower, const double* x_upper, const int& depth, const int& ilower0,          \
        const int& iupper0, const int& ilower1, const int& iupper1, const int& ilower2, const int& iupper2,            \
        const int& nugc0, const int& nugc1, const int& nugc2, const double* u, const int* indices, const double* Xshift, \
        const int& nindices, const double* X, double* V)                                                               \
    {                                                                                                                  \
        const int il[3] = { ilower0, ilower1, ilower2 }, iu[3] = { iupper0, iupper1, iupper2 };                        \
        const int ng[3] = { nugc0, nugc1, nugc2 };                                                                     \
        const ibk_array_desc d = desc(3, dx, x_lower, x_upper, depth, il, iu, ng);                                     \
        check(ibk_raw_interp_host(ctx(), KERNEL, &d, u, indices, Xshift, nindices, X, max_index(indices, nindices), V)); \
    }                                                                                                                  \
    extern "C" void lagrangian_##NAME##_spread3d_(                                                                     \
        const double* dx, const double* x_lower, const double* x_upper, const int& depth, const int* indices,          \
        const double* Xshift, const int& nindices, const double* X, const double* V, const int& ilower0,               \
        const int& iupper0, const int& ilower1, const int& iupper1, const int& ilower2, const int& iupper2,            \
        const int& nugc0, const int& nugc1, const int& nugc2, double* u)                                               \
    {                                                                                                                  \
        const int il[3] = { ilower0, ilower1, ilower2 }, iu[3] = { iupper0, iupper1, iupper2 };                        \
        const int ng[3] = { nugc0, nugc1, nugc2 };                                                                     \
        const ibk_array_desc d = desc(3, dx, x_lower, x_upper, depth, il, iu, ng);                                     \
        check(ibk_raw_spread_host(ctx(), KERNEL, &d, indices, Xshift, nindices, X, max_index(indices, nindices), V, u)); \
    }

#define IBK_SHIM_2D(NAME, KERNEL)                                                                                      \
    extern "C" void lagrangian_##NAME##_interp2d_(const double* dx, const double* x_lower, const double* x_upper,      \
                                                  const int& depth, const int& ilower0, const int& iupper0,            \
                                                  const int& ilower1, const int& iupper1, const int& nugc0,            \
                                                  const int& nugc1, const double* u, const int* indices,               \
                                                  const double* Xshift, const int& nindices, const double* X, double* V) \
    {                                                                                                                  \
        const int il[2] = { ilower0, ilower1 }, iu[2] = { iupper0, iupper1 }, ng[2] = { nugc0, nugc1 };                \
        const ibk_array_desc d = desc(2, dx, x_lower, x_upper, depth, il, iu, ng);                                     \
        check(ibk_raw_interp_host(ctx(), KERNEL, &d, u, indices, Xshift, nindices, X, max_index(indices, nindices), V)); \
    }                                                                                                                  \
    extern "C" void lagrangian_##NAME##_spread2d_(const double* dx, const double* x_lower, const double* x_upper,      \
                                                  const int& depth, const int* indices, const double* Xshift,          \
                                                  const int& nindices, const double* X, const double* V,               \
                                                  const int& ilower0, const int& iupper0, const int& ilower1,          \
                                                  const int& iupper1, const int& nugc0, const int& nugc1, double* u)   \
    {                                                                                                                  \
        const int il[2] = { ilower0, ilower1 }, iu[2] = { iupper0, iupper1 }, ng[2] = { nugc0, nugc1 };                \
        const ibk_array_desc d = desc(2, dx, x_lower, x_upper, depth, il, iu, ng);                                     \
        check(ibk_raw_spread_host(ctx(), KERNEL, &d, indices, Xshift, nindices, X, max_index(indices, nindices), V, u)); \
    }

IBK_SHIM_3D(piecewise_linear, IBK_PIECEWISE_LINEAR)
IBK_SHIM_3D(ib_4, IBK_IB_4)
IBK_SHIM_3D(ib_6, IBK_IB_6)
IBK_SHIM_3D(bspline_3, IBK_BSPLINE_3)
IBK_SHIM_3D(bspline_4, IBK_BSPLINE_4)
IBK_SHIM_3D(ib_3, IBK_IB_3)
IBK_SHIM_3D(bspline_5, IBK_BSPLINE_5)
IBK_SHIM_3D(bspline_6, IBK_BSPLINE_6)
IBK_SHIM_3D(piecewise_cubic, IBK_PIECEWISE_CUBIC)
IBK_SHIM_3D(ib_5, IBK_IB_5)
IBK_SHIM_3D(piecewise_constant, IBK_PIECEWISE_CONSTANT)
IBK_SHIM_2D(piecewise_linear, IBK_PIECEWISE_LINEAR)
IBK_SHIM_2D(ib_4, IBK_IB_4)
IBK_SHIM_2D(ib_6, IBK_IB_6)
IBK_SHIM_2D(bspline_3, IBK_BSPLINE_3)
IBK_SHIM_2D(bspline_4, IBK_BSPLINE_4)
IBK_SHIM_2D(ib_3, IBK_IB_3)
IBK_SHIM_2D(bspline_5, IBK_BSPLINE_5)
IBK_SHIM_2D(bspline_6, IBK_BSPLINE_6)
IBK_SHIM_2D(piecewise_cubic, IBK_PIECEWISE_CUBIC)
IBK_SHIM_2D(ib_5, IBK_IB_5)
IBK_SHIM_2D(piecewise_constant, IBK_PIECEWISE_CONSTANT)

// axis-dependent routines: `axis` follows `depth` (lagrangian_interaction3d.f.m4:237-243, 3510-3516; LEInteractor.cpp:299-360,
// 1021-1080)
#define IBK_SHIM_AXIS_3D(NAME, KERNEL)                                                                                      \
    extern "C" void lagrangian_##NAME##_interp3d_(                                                                     \
        const double* dx, const double* x_lower, const double* x_upper, const int& depth, const int& axis, const int& ilower0,          \
        const int& iupper0, const int& ilower1, const int& iupper1, const int& ilower2, const int& iupper2,            \
        const int& nugc0, const int& nugc1, const int& nugc2, const double* u, const int* indices, const double* Xshift, \
        const int& nindices, const double* X, double* V)                                                               \
    {                                                                                                                  \
        const int il[3] = { ilower0, ilower1, ilower2 }, iu[3] = { iupper0, iupper1, iupper2 };                        \
        const int ng[3] = { nugc0, nugc1, nugc2 };                                                                     \
        ibk_array_desc d = desc(3, dx, x_lower, x_upper, depth, il, iu, ng); d.axis = axis;                                     \
        check(ibk_raw_interp_host(ctx(), KERNEL, &d, u, indices, Xshift, nindices, X, max_index(indices, nindices), V)); \
    }                                                                                                                  \
    extern "C" void lagrangian_##NAME##_spread3d_(                                                                     \
        const double* dx, const double* x_lower, const double* x_upper, const int& depth, const int& axis, const int* indices,          \
        const double* Xshift, const int& nindices, const double* X, const double* V, const int& ilower0,               \
        const int& iupper0, const int& ilower1, const int& iupper1, const int& ilower2, const int& iupper2,            \
        const int& nugc0, const int& nugc1, const int& nugc2, double* u)                                               \
    {                                                                                                                  \
        const int il[3] = { ilower0, ilower1, ilower2 }, iu[3] = { iupper0, iupper1, iupper2 };                        \
        const int ng[3] = { nugc0, nugc1, nugc2 };                                                                     \
        ibk_array_desc d = desc(3, dx, x_lower, x_upper, depth, il, iu, ng); d.axis = axis;                                     \
        check(ibk_raw_spread_host(ctx(), KERNEL, &d, indices, Xshift, nindices, X, max_index(indices, nindices), V, u)); \
    }

#define IBK_SHIM_AXIS_2D(NAME, KERNEL)                                                                                      \
    extern "C" void lagrangian_##NAME##_interp2d_(const double* dx, const double* x_lower, const double* x_upper,      \
                                                  const int& depth, const int& axis, const int& ilower0, const int& iupper0,            \
                                                  const int& ilower1, const int& iupper1, const int& nugc0,            \
                                                  const int& nugc1, const double* u, const int* indices,               \
                                                  const double* Xshift, const int& nindices, const double* X, double* V) \
    {                                                                                                                  \
        const int il[2] = { ilower0, ilower1 }, iu[2] = { iupper0, iupper1 }, ng[2] = { nugc0, nugc1 };                \
        ibk_array_desc d = desc(2, dx, x_lower, x_upper, depth, il, iu, ng); d.axis = axis;                                     \
        check(ibk_raw_interp_host(ctx(), KERNEL, &d, u, indices, Xshift, nindices, X, max_index(indices, nindices), V)); \
    }                                                                                                                  \
    extern "C" void lagrangian_##NAME##_spread2d_(const double* dx, const double* x_lower, const double* x_upper,      \
                                                  const int& depth, const int& axis, const int* indices, const double* Xshift,          \
                                                  const int& nindices, const double* X, const double* V,               \
                                                  const int& ilower0, const int& iupper0, const int& ilower1,          \
                                                  const int& iupper1, const int& nugc0, const int& nugc1, double* u)   \
    {                                                                                                                  \
        const int il[2] = { ilower0, ilower1 }, iu[2] = { iupper0, iupper1 }, ng[2] = { nugc0, nugc1 };                \
        ibk_array_desc d = desc(2, dx, x_lower, x_upper, depth, il, iu, ng); d.axis = axis;                                     \
        check(ibk_raw_spread_host(ctx(), KERNEL, &d, indices, Xshift, nindices, X, max_index(indices, nindices), V, u)); \
    }

IBK_SHIM_3D(ib_4_w8, IBK_IB_4_W8)
IBK_SHIM_2D(ib_4_w8, IBK_IB_4_W8)
IBK_SHIM_AXIS_3D(composite_bspline_32, IBK_COMPOSITE_BSPLINE_32)
IBK_SHIM_AXIS_2D(composite_bspline_32, IBK_COMPOSITE_BSPLINE_32)
IBK_SHIM_AXIS_3D(composite_bspline_23, IBK_COMPOSITE_BSPLINE_23)
IBK_SHIM_AXIS_2D(composite_bspline_23, IBK_COMPOSITE_BSPLINE_23)
IBK_SHIM_AXIS_3D(composite_bspline_43, IBK_COMPOSITE_BSPLINE_43)
IBK_SHIM_AXIS_2D(composite_bspline_43, IBK_COMPOSITE_BSPLINE_43)
IBK_SHIM_AXIS_3D(composite_bspline_34, IBK_COMPOSITE_BSPLINE_34)
IBK_SHIM_AXIS_2D(composite_bspline_34, IBK_COMPOSITE_BSPLINE_34)
IBK_SHIM_AXIS_3D(composite_bspline_54, IBK_COMPOSITE_BSPLINE_54)
IBK_SHIM_AXIS_2D(composite_bspline_54, IBK_COMPOSITE_BSPLINE_54)
IBK_SHIM_AXIS_3D(composite_bspline_45, IBK_COMPOSITE_BSPLINE_45)
IBK_SHIM_AXIS_2D(composite_bspline_45, IBK_COMPOSITE_BSPLINE_45)
IBK_SHIM_AXIS_3D(composite_bspline_65, IBK_COMPOSITE_BSPLINE_65)
IBK_SHIM_AXIS_2D(composite_bspline_65, IBK_COMPOSITE_BSPLINE_65)
IBK_SHIM_AXIS_3D(composite_bspline_56, IBK_COMPOSITE_BSPLINE_56)
IBK_SHIM_AXIS_2D(composite_bspline_56, IBK_COMPOSITE_BSPLINE_56)
IBK_SHIM_AXIS_3D(discontinuous_linear, IBK_DISCONTINUOUS_LINEAR)
IBK_SHIM_AXIS_2D(discontinuous_linear, IBK_DISCONTINUOUS_LINEAR)
