// LDataManagerB200.h -- the LDataManager::spread / interp entry points (seam B2 of SURVEY.md 8(b)) over the device-resident
// level: ibtk/src/lagrangian/LDataManager.cpp:551-562 (spread with explicit F / X data, kernel function, boundary op and
// prolongation schedules), :416-447 (the overload that first forms F * ds from node weights), :698-706 (interp).
// F_data / X_data are LDataB200 views of marker columns of the SAME context; a column other than the working ones is
// brought into place on the device (ibk_markers_lincomb), nothing crosses PCIe.  coarsest_ln / finest_ln select levels
// of the hierarchy in the reference; this object holds one level (the finest, where the structure lives): any other
// range is refused.
#pragma once
#include "IBMethodB200.h"
#include "LDataB200.h"

namespace IBTK_B200
{
using namespace SAMRAI_standin;

class LDataManagerB200
{
public:
    static constexpr int invalid_level_number = -1;
    explicit LDataManagerB200(IBAMR_B200::IBMethodB200& level, int level_number = 0) : d_ib(level), d_ln(level_number)
    {
    }
    // The next coarser level of the hierarchy (resident on the same device, its own context) and the refinement ratio to it.
    // With it, spread() first prolongs the coarser level's f onto this level when f_prolongation_scheds[ln] is set
    // (LDataManager.cpp:611-614) and interp() first coarsens this level's u onto it when f_synch_scheds[ln] is set
    // (:728-734): ibk_amr_refine_side / ibk_amr_coarsen_side.
    void setCoarserLevel(IBAMR_B200::IBMethodB200* coarser, const int* ratio)
    {
        d_coarser = coarser;
        for (int d = 0; d < 3; ++d) d_ratio[d] = ratio && coarser ? ratio[d] : 1;
    }
    // LDataManager::spread(f_data_idx, F_data, X_data, spread_kernel_fcn, f_phys_bdry_op, f_prolongation_scheds,
    //                      fill_data_time, F_data_ghost_node_update, X_data_ghost_node_update, coarsest_ln, finest_ln)
    void spread(int f_data_idx, std::vector<Pointer<LDataB200>>& F_data, std::vector<Pointer<LDataB200>>& X_data,
                const std::string& spread_kernel_fcn, RobinPhysBdryPatchStrategy* f_phys_bdry_op,
                const std::vector<Pointer<RefineSchedule>>& f_prolongation_scheds, double fill_data_time,
                bool /*F_data_ghost_node_update*/ = true, bool /*X_data_ghost_node_update*/ = true, int coarsest_ln = invalid_level_number,
                int finest_ln = invalid_level_number)
    {
        check_levels(coarsest_ln, finest_ln);
        prolong(f_prolongation_scheds);
        stage(*F_data.at(d_ln), IBK_COL_F);
        stage(*X_data.at(d_ln), IBK_COL_X);
        d_ib.setKernels(spread_kernel_fcn, spread_kernel_fcn);
        d_ib.spreadForce(f_data_idx, f_phys_bdry_op, f_prolongation_scheds, fill_data_time);
    }
    // the overload with node weights: F * ds first (LDataManager.cpp:416-447), on the device, into the aux column
    void spread(int f_data_idx, std::vector<Pointer<LDataB200>>& F_data, std::vector<Pointer<LDataB200>>& X_data,
                const std::vector<double>& ds /* [n] per-node weights, Lagrangian order */, const std::string& spread_kernel_fcn,
                RobinPhysBdryPatchStrategy* f_phys_bdry_op, const std::vector<Pointer<RefineSchedule>>& f_prolongation_scheds,
                double fill_data_time)
    {
        prolong(f_prolongation_scheds);
        stage(*X_data.at(d_ln), IBK_COL_X);
        if (ibk_markers_scale_rows(d_ib.ctx(), IBK_COL_F, F_data.at(d_ln)->column(), ds.data()) != IBK_OK)
            throw std::runtime_error(std::string("LDataManagerB200::spread: ") + ibk_last_error(d_ib.ctx()));
        d_ib.setKernels(spread_kernel_fcn, spread_kernel_fcn);
        d_ib.spreadForce(f_data_idx, f_phys_bdry_op, f_prolongation_scheds, fill_data_time);
    }
    // LDataManager::interp(f_data_idx, F_data, X_data, f_synch_scheds, f_ghost_fill_scheds, fill_data_time, coarsest_ln, finest_ln)
    void interp(int f_data_idx, std::vector<Pointer<LDataB200>>& F_data, std::vector<Pointer<LDataB200>>& X_data,
                const std::vector<Pointer<CoarsenSchedule>>& f_synch_scheds, const std::vector<Pointer<RefineSchedule>>& f_ghost_fill_scheds,
                double fill_data_time, int coarsest_ln = invalid_level_number, int finest_ln = invalid_level_number)
    {
        check_levels(coarsest_ln, finest_ln);
        if (d_coarser && d_ln < (int)f_synch_scheds.size() && f_synch_scheds[d_ln] &&
            ibk_amr_coarsen_side(d_coarser->ctx(), d_ib.ctx(), 0, d_ratio, nullptr) != IBK_OK)
            throw std::runtime_error(std::string("LDataManagerB200::interp: ") + ibk_last_error(d_coarser->ctx()));
        stage(*X_data.at(d_ln), IBK_COL_X);
        d_ib.interpolateVelocity(f_data_idx, f_synch_scheds, f_ghost_fill_scheds, fill_data_time);
        // the result lands in the U column; hand it to the caller's LData
        LDataB200& out = *F_data.at(d_ln);
        if (out.column() != IBK_COL_U && ibk_markers_lincomb(d_ib.ctx(), out.column(), 1.0, IBK_COL_U, 0.0, IBK_COL_U) != IBK_OK)
            throw std::runtime_error(std::string("LDataManagerB200::interp: ") + ibk_last_error(d_ib.ctx()));
        out.markDeviceModified();
    }

private:
    void check_levels(int coarsest_ln, int finest_ln) const
    {
        const int c = coarsest_ln == invalid_level_number ? d_ln : coarsest_ln, f = finest_ln == invalid_level_number ? d_ln : finest_ln;
        if (c != d_ln || f != d_ln) throw std::runtime_error("LDataManagerB200: this object holds level " + std::to_string(d_ln) + " only");
    }
    void prolong(const std::vector<Pointer<RefineSchedule>>& f_prolongation_scheds)
    {
        if (d_coarser && d_ln < (int)f_prolongation_scheds.size() && f_prolongation_scheds[d_ln] &&
            ibk_amr_refine_side(d_ib.ctx(), d_coarser->ctx(), 1, d_ratio, nullptr) != IBK_OK)
            throw std::runtime_error(std::string("LDataManagerB200::spread: ") + ibk_last_error(d_ib.ctx()));
    }
    void stage(LDataB200& data, int working_column)
    {
        data.restoreArrays(); // a host array handed out and modified goes to the device first
        if (data.column() == working_column) return;
        if (ibk_markers_lincomb(d_ib.ctx(), working_column, 1.0, data.column(), 0.0, data.column()) != IBK_OK)
            throw std::runtime_error(std::string("LDataManagerB200: ") + ibk_last_error(d_ib.ctx()));
    }
    IBAMR_B200::IBMethodB200& d_ib;
    int d_ln;
    IBAMR_B200::IBMethodB200* d_coarser = nullptr;
    int d_ratio[3] = { 1, 1, 1 };
};
} // namespace IBTK_B200
