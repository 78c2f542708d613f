// IBMethodB200.h -- the IBStrategy-shaped object an IBHierarchyIntegrator would hold for the hot path
// (include/ibamr/IBStrategy.h:276-280, 338-342, 455, 464; overridden by src/IB/IBMethod.cpp:672-694,
// 972-995, 1494-1557).  The PatchHierarchy / data-index arguments of the reference are replaced by the
// level registered at construction (the patches THIS process owns) and by the device-resident u / f;
// LData X / U / F are exchanged with the host in Lagrangian order (LData AoS layout, LData.h:351-367).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ibk.h"
#include "samrai_standins.h"

namespace IBAMR_B200
{
using namespace SAMRAI_standin;

class IBMethodB200
{
public:
    struct LevelSpec
    {
        Box domain_box;                        // level index space of the physical domain
        std::array<double, NDIM> x_lower{}, x_upper{};
        std::array<int, NDIM> periodic{};
        std::vector<Box> patch_boxes;          // patches owned by this process
    };

    IBMethodB200(const LevelSpec& level, const std::string& kernel_fcn = "IB_4", int device = 0, int ghost_width = -1)
        : d_interp_kernel_fcn(kernel_fcn), d_spread_kernel_fcn(kernel_fcn)
    {
        if (ibk_ctx_create(device, &d_ctx) != IBK_OK) throw std::runtime_error("IBMethodB200: no CUDA device");
        d_ghosts = ghost_width >= 0 ? ghost_width : getMinimumGhostCellWidth()(0);
        ibk_level_desc ld{};
        ld.ndim = NDIM;
        ld.n_patches = (int)level.patch_boxes.size();
        std::vector<int> lo, hi;
        for (const Box& b : level.patch_boxes)
            for (int d = 0; d < NDIM; ++d)
            {
                lo.push_back(b.lo(d));
                hi.push_back(b.hi(d));
            }
        for (int d = 0; d < NDIM; ++d)
        {
            ld.domain_lower[d] = level.domain_box.lo(d);
            ld.domain_upper[d] = level.domain_box.hi(d);
            ld.x_lower[d] = level.x_lower[d];
            ld.x_upper[d] = level.x_upper[d];
            ld.periodic[d] = level.periodic[d];
            ld.gcw[d] = d_ghosts;
        }
        ld.patch_lower = lo.data();
        ld.patch_upper = hi.data();
        check(ibk_level_create(d_ctx, &ld));
    }
    ~IBMethodB200()
    {
        if (d_ctx) ibk_ctx_destroy(d_ctx);
    }
    IBMethodB200(const IBMethodB200&) = delete;
    IBMethodB200& operator=(const IBMethodB200&) = delete;

    // IBStrategy::getMinimumGhostCellWidth (IBMethod.cpp:266-270)
    IntVector getMinimumGhostCellWidth() const
    {
        const int a = ibk_get_minimum_ghost_width(d_interp_kernel_fcn.c_str());
        const int b = ibk_get_minimum_ghost_width(d_spread_kernel_fcn.c_str());
        return IntVector(a > b ? a : b);
    }

    // LData("X") set-up + Lagrangian numbering (LDataManager::initializeLevelData role)
    void setPositions(const std::vector<double>& X)
    {
        d_n = (int)(X.size() / NDIM);
        check(ibk_markers_set_positions(d_ctx, X.data(), d_n));
    }
    void setForce(const std::vector<double>& F)
    {
        check(ibk_markers_upload(d_ctx, 2, F.data()));
    }
    void getVelocity(std::vector<double>& U)
    {
        U.resize((size_t)d_n * NDIM);
        check(ibk_markers_download(d_ctx, 1, U.data()));
    }
    // u_data_idx / f_data_idx of the reference: SideData arrays of patch p
    void setEulerianVelocity(int p, const SideData& u)
    {
        for (int a = 0; a < NDIM; ++a) check(ibk_grid_upload(d_ctx, 0, p, a, u.getPointer(a)));
    }
    void setEulerianForce(int p, const SideData& f)
    {
        for (int a = 0; a < NDIM; ++a) check(ibk_grid_upload(d_ctx, 1, p, a, f.getPointer(a)));
    }
    void getEulerianForce(int p, SideData& f)
    {
        for (int a = 0; a < NDIM; ++a) check(ibk_grid_download(d_ctx, 1, p, a, f.getPointer(a)));
    }

    // IBStrategy::beginDataRedistribution / endDataRedistribution (IBStrategy.h:455, 464)
    void beginDataRedistribution()
    {
        check(ibk_rebin(d_ctx, d_error_if_points_leave_domain ? 1 : 0));
    }
    void endDataRedistribution()
    {
    }

    // IBStrategy::interpolateVelocity(u_data_idx, u_synch_scheds, u_ghost_fill_scheds, data_time)
    void interpolateVelocity(double /*data_time*/ = 0.0)
    {
        check(ibk_interpolate_velocity(d_ctx, d_interp_kernel_fcn.c_str(), /*fill_halo*/ 1));
    }
    // IBStrategy::spreadForce(f_data_idx, f_phys_bdry_op, f_prolongation_scheds, data_time)
    void spreadForce(double /*data_time*/ = 0.0)
    {
        check(ibk_spread_force(d_ctx, d_spread_kernel_fcn.c_str(), /*accumulate_halo*/ 1));
    }

    ibk_ctx* ctx()
    {
        return d_ctx;
    }
    bool d_error_if_points_leave_domain = false; // IBMethod.cpp:2060

private:
    void check(int rc)
    {
        if (rc != IBK_OK) throw std::runtime_error(std::string("IBMethodB200: ") + ibk_last_error(d_ctx));
    }
    ibk_ctx* d_ctx = nullptr;
    std::string d_interp_kernel_fcn, d_spread_kernel_fcn; // IBMethod.h:602 default "IB_4"
    int d_ghosts = 0;
    int d_n = 0;
};
} // namespace IBAMR_B200
