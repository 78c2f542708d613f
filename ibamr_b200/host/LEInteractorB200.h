// LEInteractorB200.h -- host-side C++ mirror of IBTK::LEInteractor for the hot path
// (ibtk/include/ibtk/LEInteractor.h:75): same static method names, argument order and meaning, and the
// same error conditions (thrown as std::runtime_error where the reference calls TBOX_ERROR), forwarding
// to the C ABI of libibk.so.  Header-only; link with -libk.
//
//   getStencilSize / getMinimumGhostWidth / isKnownKernel     LEInteractor.h:99-117
//   interpolate(Q, Q_size, Q_depth, X, X_size, X_depth, SideData|CellData, patch, box, fcn)
//                                                             LEInteractor.h:566-575 (.cpp:3045-3127, 2805-2863)
//   spread(SideData|CellData, Q, Q_size, Q_depth, X, X_size, X_depth, patch, box, fcn)
//                                                             LEInteractor.h:1132-1141 (.cpp:4188-4263, 3951-4009)
//   interpolate / spread with (local_indices, periodic_shifts): the index-set overloads
//                                                             LEInteractor.h:184-192, 704-712 with the lists of
//                                                             LIndexSetData (LIndexSetData.h:79-145)
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ibk.h"
#include "samrai_standins.h"

namespace IBTK_B200
{
using namespace SAMRAI_standin;

class LEInteractor
{
public:
    // One context per process/device, created on first use (LEInteractor is all-static in the reference).
    static ibk_ctx* context(int device = 0)
    {
        static ibk_ctx* ctx = nullptr;
        if (!ctx)
        {
            const int rc = ibk_ctx_create(device, &ctx);
            if (rc != IBK_OK) throw std::runtime_error("LEInteractor: no CUDA device (libibk has no CPU fallback)");
        }
        return ctx;
    }

    static bool isKnownKernel(const std::string& kernel_fcn)
    {
        return ibk_is_known_kernel(kernel_fcn.c_str()) != 0;
    }
    static int getStencilSize(const std::string& kernel_fcn)
    {
        const int r = ibk_get_stencil_size(kernel_fcn.c_str());
        if (r < 0) throw std::runtime_error("LEInteractor::getStencilSize()\n  Unknown kernel function " + kernel_fcn);
        return r;
    }
    static int getMinimumGhostWidth(const std::string& kernel_fcn)
    {
        const int r = ibk_get_minimum_ghost_width(kernel_fcn.c_str());
        if (r < 0) throw std::runtime_error("LEInteractor::getMinimumGhostWidth()\n  Unknown kernel function " + kernel_fcn);
        return r;
    }

    // ---- position-only interpolation, SideData
    static void interpolate(double* const Q_data, const int Q_size, const int Q_depth, const double* const X_data,
                            const int X_size, const int X_depth, const Pointer<SideData> q_data, const Pointer<Patch> patch,
                            const Box& interp_box, const std::string& interp_fcn = "IB_4")
    {
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        const double* q[NDIM];
        for (int a = 0; a < NDIM; ++a) q[a] = q_data->getPointer(a);
        check(ibk_side_interpolate_host(context(), interp_fcn.c_str(), &pd, q, q_data->getDepth(), interp_box.lo.v.data(),
                                        interp_box.hi.v.data(), X_data, X_size, X_depth, Q_data, Q_size, Q_depth),
              "LEInteractor::interpolate()");
    }
    // ---- position-only interpolation, CellData
    static void interpolate(double* const Q_data, const int Q_size, const int Q_depth, const double* const X_data,
                            const int X_size, const int X_depth, const Pointer<CellData> q_data, const Pointer<Patch> patch,
                            const Box& interp_box, const std::string& interp_fcn = "IB_4")
    {
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        check(ibk_cell_interpolate_host(context(), interp_fcn.c_str(), &pd, q_data->getPointer(), q_data->getDepth(),
                                        interp_box.lo.v.data(), interp_box.hi.v.data(), X_data, X_size, X_depth, Q_data, Q_size,
                                        Q_depth),
              "LEInteractor::interpolate()");
    }
    // std::vector conveniences (LEInteractor.h:577-628)
    template <class DataT>
    static void interpolate(std::vector<double>& Q_data, const int Q_depth, const std::vector<double>& X_data, const int X_depth,
                            const Pointer<DataT> q_data, const Pointer<Patch> patch, const Box& interp_box,
                            const std::string& interp_fcn = "IB_4")
    {
        interpolate(Q_data.data(), (int)Q_data.size(), Q_depth, X_data.data(), (int)X_data.size(), X_depth, q_data, patch,
                    interp_box, interp_fcn);
    }

    // ---- position-only spreading
    static void spread(Pointer<SideData> q_data, const double* const Q_data, const int Q_size, const int Q_depth,
                       const double* const X_data, const int X_size, const int X_depth, const Pointer<Patch> patch,
                       const Box& spread_box, const std::string& spread_fcn = "IB_4")
    {
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        double* q[NDIM];
        for (int a = 0; a < NDIM; ++a) q[a] = q_data->getPointer(a);
        check(ibk_side_spread_host(context(), spread_fcn.c_str(), &pd, q, q_data->getDepth(), spread_box.lo.v.data(),
                                   spread_box.hi.v.data(), X_data, X_size, X_depth, Q_data, Q_size, Q_depth),
              "LEInteractor::spread()");
    }
    static void spread(Pointer<CellData> q_data, const double* const Q_data, const int Q_size, const int Q_depth,
                       const double* const X_data, const int X_size, const int X_depth, const Pointer<Patch> patch,
                       const Box& spread_box, const std::string& spread_fcn = "IB_4")
    {
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        check(ibk_cell_spread_host(context(), spread_fcn.c_str(), &pd, q_data->getPointer(), q_data->getDepth(),
                                   spread_box.lo.v.data(), spread_box.hi.v.data(), X_data, X_size, X_depth, Q_data, Q_size, Q_depth),
              "LEInteractor::spread()");
    }
    template <class DataT>
    static void spread(Pointer<DataT> q_data, const std::vector<double>& Q_data, const int Q_depth, const std::vector<double>& X_data,
                       const int X_depth, const Pointer<Patch> patch, const Box& spread_box, const std::string& spread_fcn = "IB_4")
    {
        spread(q_data, Q_data.data(), (int)Q_data.size(), Q_depth, X_data.data(), (int)X_data.size(), X_depth, patch, spread_box,
               spread_fcn);
    }

    // ---- index-set forms: the flat lists LIndexSetData caches (local PETSc indices + periodic shifts)
    static void interpolate(double* const Q_data, const int Q_depth, const double* const X_data, const int X_depth,
                            const std::vector<int>& local_indices, const std::vector<double>& periodic_shifts, const int n_markers,
                            const Pointer<SideData> q_data, const Pointer<Patch> patch, const std::string& interp_fcn = "IB_4")
    {
        if (Q_depth != NDIM || X_depth != NDIM || q_data->getDepth() != 1)
            throw std::runtime_error("LEInteractor::interpolate():\n  side-centered interpolation requires vector-valued data.\n");
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        const double* q[NDIM];
        for (int a = 0; a < NDIM; ++a) q[a] = q_data->getPointer(a);
        check(ibk_side_interpolate_indexed_host(context(), interp_fcn.c_str(), &pd, q, local_indices.data(),
                                                periodic_shifts.empty() ? nullptr : periodic_shifts.data(),
                                                (int)local_indices.size(), X_data, n_markers, Q_data),
              "LEInteractor::interpolate()");
    }
    static void spread(Pointer<SideData> q_data, const double* const Q_data, const int Q_depth, const double* const X_data,
                       const int X_depth, const std::vector<int>& local_indices, const std::vector<double>& periodic_shifts,
                       const int n_markers, const Pointer<Patch> patch, const std::string& spread_fcn = "IB_4")
    {
        if (Q_depth != NDIM || X_depth != NDIM || q_data->getDepth() != 1)
            throw std::runtime_error("LEInteractor::spread():\n  side-centered spreading requires vector-valued data.\n");
        const ibk_patch_desc pd = patch_desc(*patch, q_data->getGhostCellWidth());
        double* q[NDIM];
        for (int a = 0; a < NDIM; ++a) q[a] = q_data->getPointer(a);
        check(ibk_side_spread_indexed_host(context(), spread_fcn.c_str(), &pd, q, local_indices.data(),
                                           periodic_shifts.empty() ? nullptr : periodic_shifts.data(), (int)local_indices.size(),
                                           X_data, n_markers, Q_data),
              "LEInteractor::spread()");
    }

    static ibk_patch_desc patch_desc(const Patch& patch, const IntVector& gcw)
    {
        ibk_patch_desc pd{};
        pd.ndim = NDIM;
        const auto& g = *patch.getPatchGeometry();
        for (int d = 0; d < NDIM; ++d)
        {
            pd.lower[d] = patch.box.lo(d);
            pd.upper[d] = patch.box.hi(d);
            pd.gcw[d] = gcw(d);
            pd.x_lower[d] = g.x_lower[d];
            pd.x_upper[d] = g.x_upper[d];
            pd.dx[d] = g.dx[d];
        }
        pd.touches_physical_bdry = g.touches_regular_bdry ? 1 : 0;
        return pd;
    }

private:
    static void check(int rc, const char* where)
    {
        if (rc != IBK_OK) throw std::runtime_error(std::string(where) + ":\n  " + ibk_last_error(context()));
    }
};
} // namespace IBTK_B200
