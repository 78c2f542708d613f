// IBMethodB200.h -- the IBStrategy-shaped object an IBHierarchyIntegrator would hold for the hot path
// (include/ibamr/IBStrategy.h:276-280, 338-342, 455, 464; overridden by src/IB/IBMethod.cpp:672-694,
// 972-995, 1494-1557).  interpolateVelocity / spreadForce carry the reference's signatures; the PatchHierarchy behind the
// data indices is the level registered at construction (the patches THIS process owns) plus the host SideData bound to
// an index with registerPatchData; the schedule arguments are accepted and ignored (the library does their work).
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

    // ---- patch data indices: the reference passes u_data_idx / f_data_idx into the PatchHierarchy; here a data index is
    // bound to the host SideData of each local patch (the fluid solver's arrays).  A bound index is uploaded before and,
    // for f, downloaded after the operation; an unbound index means "the resident u / f of the library" (device-resident
    // fluid data, or data moved with setEulerianVelocity / getEulerianForce).
    void registerPatchData(int data_idx, int patch, SideData* data)
    {
        if ((int)d_patch_data.size() <= data_idx) d_patch_data.resize(data_idx + 1);
        if ((int)d_patch_data[data_idx].size() <= patch) d_patch_data[data_idx].resize(patch + 1, nullptr);
        d_patch_data[data_idx][patch] = data;
    }

    // ---- more than one rank (LDataManager.cpp:597-620, 744): the communicator of this rank's context and the global patch
    // list; from then on interpolateVelocity / spreadForce run the inter-rank ghost fill / ghost accumulation with the
    // messages in flight while the tiles that do not touch the exchanged regions are processed.
    void initCommunicator(const void* nccl_unique_id_128, int rank, int nranks)
    {
        check(ibk_comm_init(d_ctx, nccl_unique_id_128, rank, nranks));
    }
    static void initLoopbackCommunicator(const std::vector<IBMethodB200*>& ranks) // the ranks are objects of ONE process
    {
        std::vector<ibk_ctx*> c;
        for (IBMethodB200* r : ranks) c.push_back(r->d_ctx);
        if (ibk_comm_init_loopback(c.data(), (int)c.size()) != IBK_OK) throw std::runtime_error("IBMethodB200: loopback communicator");
    }
    void setGlobalPatches(const std::vector<Box>& boxes, const std::vector<int>& ranks)
    {
        std::vector<int> lo, hi;
        for (const Box& b : boxes)
            for (int d = 0; d < NDIM; ++d)
            {
                lo.push_back(b.lo(d));
                hi.push_back(b.hi(d));
            }
        check(ibk_comm_set_patches(d_ctx, (int)boxes.size(), lo.data(), hi.data(), ranks.data()));
        d_multi_rank = true;
    }

    // IBStrategy::interpolateVelocity (IBStrategy.h:276-280; IBMethod.cpp:672-694).  The schedules are stand-ins: the ghost
    // fill they perform is done by the library.  begin... / finish... are the two halves around the point where the messages
    // of the other ranks must have been posted: one process per rank calls the whole; a process that holds several ranks
    // (loopback communicator) calls begin on every rank, then finish on every rank.
    void interpolateVelocity(int u_data_idx, const std::vector<Pointer<CoarsenSchedule>>& /*u_synch_scheds*/,
                             const std::vector<Pointer<RefineSchedule>>& /*u_ghost_fill_scheds*/, double /*data_time*/)
    {
        beginInterpolateVelocity(u_data_idx);
        finishInterpolateVelocity();
    }
    void beginInterpolateVelocity(int u_data_idx)
    {
        upload(u_data_idx, 0);
        if (!d_multi_rank)
        {
            check(ibk_interpolate_velocity(d_ctx, d_interp_kernel_fcn.c_str(), /*fill_halo*/ 1));
            return;
        }
        check(ibk_halo_fill_post(d_ctx));
        check(ibk_halo_local(d_ctx, 0));
        check(ibk_interpolate_velocity_part(d_ctx, d_interp_kernel_fcn.c_str(), 1)); // interior tiles read no ghost cell
    }
    void finishInterpolateVelocity()
    {
        if (!d_multi_rank) return;
        check(ibk_halo_fill_finish(d_ctx));
        check(ibk_interpolate_velocity_part(d_ctx, d_interp_kernel_fcn.c_str(), 2));
    }
    // IBStrategy::spreadForce (IBStrategy.h:338-342; IBMethod.cpp:972-995): f += S[F].  f_phys_bdry_op and the prolongation
    // schedules are stand-ins (see samrai_standins.h).
    void spreadForce(int f_data_idx, RobinPhysBdryPatchStrategy* /*f_phys_bdry_op*/,
                     const std::vector<Pointer<RefineSchedule>>& /*f_prolongation_scheds*/, double /*data_time*/)
    {
        beginSpreadForce(f_data_idx);
        finishSpreadForce(f_data_idx);
    }
    void beginSpreadForce(int f_data_idx)
    {
        upload(f_data_idx, 1);
        if (!d_multi_rank)
        {
            check(ibk_spread_force(d_ctx, d_spread_kernel_fcn.c_str(), /*accumulate_halo*/ 1));
            return;
        }
        check(ibk_spread_begin(d_ctx));
        check(ibk_spread_force_part(d_ctx, d_spread_kernel_fcn.c_str(), 2)); // boundary tiles: everything the neighbours need
        check(ibk_halo_accumulate_post(d_ctx));
        check(ibk_spread_force_part(d_ctx, d_spread_kernel_fcn.c_str(), 1)); // interior tiles, the messages in flight
        check(ibk_halo_local(d_ctx, 1));
    }
    void finishSpreadForce(int f_data_idx)
    {
        if (d_multi_rank)
        {
            check(ibk_halo_accumulate_finish(d_ctx));
            check(ibk_spread_end(d_ctx));
        }
        download(f_data_idx, 1);
    }
    // the resident-data forms (device-resident u / f, no data index)
    void interpolateVelocity(double data_time = 0.0)
    {
        interpolateVelocity(-1, {}, {}, data_time);
    }
    void spreadForce(double data_time = 0.0)
    {
        spreadForce(-1, nullptr, {}, data_time);
    }

    // ---- N1: the steps either side of the path, with X, U, F resident on the device ----------------------
    // IBMethod::preprocessIntegrateData keeps X_current; the working column IBK_COL_X plays X_LE / half data.
    void preprocessIntegrateData()
    {
        check(ibk_markers_lincomb(d_ctx, IBK_COL_X_CURRENT, 1.0, IBK_COL_X, 0.0, IBK_COL_X));
    }
    // IBMethod::forwardEulerStep (IBMethod.cpp:714-738) + reinitMidpointData (:1900-1912)
    void forwardEulerStep(double current_time, double new_time)
    {
        const double dt = new_time - current_time;
        check(ibk_markers_lincomb(d_ctx, IBK_COL_X_NEW, 1.0, IBK_COL_X_CURRENT, dt, IBK_COL_U));
        check(ibk_markers_lincomb(d_ctx, IBK_COL_X, 0.5, IBK_COL_X_CURRENT, 0.5, IBK_COL_X_NEW));
    }
    // IBMethod::midpointStep (IBMethod.cpp:768-792): U holds the half-time velocity
    void midpointStep(double current_time, double new_time)
    {
        forwardEulerStep(current_time, new_time);
    }
    // IBMethod::postprocessIntegrateData: X_new becomes the working (and, at the next preprocess, the current) data
    void postprocessIntegrateData()
    {
        check(ibk_markers_lincomb(d_ctx, IBK_COL_X, 1.0, IBK_COL_X_NEW, 0.0, IBK_COL_X_NEW));
    }
    // IBStandardForceGen::initializeLevelData roles (IBStandardForceGen.cpp:715-811, 933-1035, 1150-1199)
    void registerSprings(const std::vector<int>& master, const std::vector<int>& slave, const std::vector<double>& kappa,
                         const std::vector<double>& rest_length)
    {
        check(ibk_force_set_springs(d_ctx, (int)master.size(), master.data(), slave.data(), kappa.data(), rest_length.data()));
    }
    void registerBeams(const std::vector<int>& curr, const std::vector<int>& next, const std::vector<int>& prev,
                       const std::vector<double>& rigidity, const std::vector<double>& curvature)
    {
        check(ibk_force_set_beams(d_ctx, (int)curr.size(), curr.data(), next.data(), prev.data(), rigidity.data(),
                                  curvature.empty() ? nullptr : curvature.data()));
    }
    void registerTargetPoints(const std::vector<int>& idx, const std::vector<double>& kappa, const std::vector<double>& eta,
                              const std::vector<double>& X0)
    {
        check(ibk_force_set_target_points(d_ctx, (int)idx.size(), idx.data(), kappa.data(), eta.empty() ? nullptr : eta.data(),
                                          X0.data()));
    }
    // IBMethod::computeLagrangianForce(data_time) (IBMethod.cpp:834-858)
    void computeLagrangianForce(double /*data_time*/ = 0.0)
    {
        check(ibk_compute_lagrangian_force(d_ctx, IBK_COL_X, IBK_COL_U, IBK_COL_F));
    }
    // IBMethod::resetAnchorPointValues (IBMethod.cpp:1915-1943)
    void resetAnchorPointValues(int column, const std::vector<int>& anchor_idx)
    {
        check(ibk_markers_zero_rows(d_ctx, column, anchor_idx.data(), (int)anchor_idx.size()));
    }
    void getColumn(int column, std::vector<double>& out)
    {
        out.assign((size_t)ibk_markers_count(d_ctx) * NDIM, 0.0);
        check(ibk_markers_download(d_ctx, column, out.data()));
    }

    ibk_ctx* ctx()
    {
        return d_ctx;
    }
    void setKernels(const std::string& interp_fcn, const std::string& spread_fcn)
    {
        d_interp_kernel_fcn = interp_fcn;
        d_spread_kernel_fcn = spread_fcn;
    }
    bool d_error_if_points_leave_domain = false; // IBMethod.cpp:2060

private:
    void check(int rc)
    {
        if (rc != IBK_OK) throw std::runtime_error(std::string("IBMethodB200: ") + ibk_last_error(d_ctx));
    }
    void upload(int data_idx, int which)
    {
        if (data_idx < 0 || data_idx >= (int)d_patch_data.size()) return;
        for (size_t p = 0; p < d_patch_data[data_idx].size(); ++p)
            if (SideData* sd = d_patch_data[data_idx][p])
                for (int a = 0; a < NDIM; ++a) check(ibk_grid_upload(d_ctx, which, (int)p, a, sd->getPointer(a)));
    }
    void download(int data_idx, int which)
    {
        if (data_idx < 0 || data_idx >= (int)d_patch_data.size()) return;
        for (size_t p = 0; p < d_patch_data[data_idx].size(); ++p)
            if (SideData* sd = d_patch_data[data_idx][p])
                for (int a = 0; a < NDIM; ++a) check(ibk_grid_download(d_ctx, which, (int)p, a, sd->getPointer(a)));
    }
    std::vector<std::vector<SideData*>> d_patch_data; // [data_idx][local patch]
    bool d_multi_rank = false;
    ibk_ctx* d_ctx = nullptr;
    std::string d_interp_kernel_fcn, d_spread_kernel_fcn; // IBMethod.h:602 default "IB_4"
    int d_ghosts = 0;
    int d_n = 0;
};
} // namespace IBAMR_B200
