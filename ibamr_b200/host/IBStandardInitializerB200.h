// IBStandardInitializerB200.h -- host mirror of IBAMR::IBStandardInitializer's file reading
// (src/IB/IBStandardInitializer.cpp:131-182 init; readers :184-294, 297-528, 766-1002, 1322-1517, 1520-1643) over
// libibk.so's ibk_io_* entry points, for ONE level.  Structures are given by base filename; <base>.vertex is
// required, .spring/.beam/.target/.anchor are optional; vertex numbers of structure j are offset by the vertex
// counts of the structures before it (:203-210).  Errors throw (TBOX_ERROR in the reference).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ibk.h"

namespace IBAMR_B200
{
class IBStandardInitializerB200
{
public:
    IBStandardInitializerB200(int ndim, const std::vector<std::string>& base_filenames) : d_ndim(ndim)
    {
        int offset = 0, n = 0;
        for (const std::string& base : base_filenames)
        {
            const std::string v = base + ".vertex", s = base + ".spring", b = base + ".beam", t = base + ".target", a = base + ".anchor";
            check(ibk_io_read_vertex_file(v.c_str(), ndim, nullptr, 0, &n));
            const int nv = n;
            const size_t x0 = X.size();
            X.resize(x0 + (size_t)nv * ndim);
            check(ibk_io_read_vertex_file(v.c_str(), ndim, X.data() + x0, nv, &n));

            check(ibk_io_read_spring_file(s.c_str(), nv, offset, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &n));
            const size_t s0 = spring_master.size();
            grow(n, spring_master, spring_slave, spring_fcn);
            spring_kappa.resize(s0 + n);
            spring_rest.resize(s0 + n);
            if (n)
                check(ibk_io_read_spring_file(s.c_str(), nv, offset, spring_master.data() + s0, spring_slave.data() + s0,
                                              spring_kappa.data() + s0, spring_rest.data() + s0, spring_fcn.data() + s0, n, &n));

            check(ibk_io_read_beam_file(b.c_str(), nv, offset, ndim, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &n));
            const size_t b0 = beam_curr.size();
            grow(n, beam_prev, beam_curr, beam_next);
            beam_rigidity.resize(b0 + n);
            beam_curvature.resize((b0 + n) * ndim);
            if (n)
                check(ibk_io_read_beam_file(b.c_str(), nv, offset, ndim, beam_prev.data() + b0, beam_curr.data() + b0,
                                            beam_next.data() + b0, beam_rigidity.data() + b0, beam_curvature.data() + b0 * ndim, n, &n));

            check(ibk_io_read_target_file(t.c_str(), nv, offset, nullptr, nullptr, nullptr, 0, &n));
            const size_t t0 = target_idx.size();
            target_idx.resize(t0 + n);
            target_kappa.resize(t0 + n);
            target_eta.resize(t0 + n);
            if (n)
                check(ibk_io_read_target_file(t.c_str(), nv, offset, target_idx.data() + t0, target_kappa.data() + t0,
                                              target_eta.data() + t0, n, &n));

            check(ibk_io_read_anchor_file(a.c_str(), nv, offset, nullptr, 0, &n));
            const size_t a0 = anchor_idx.size();
            anchor_idx.resize(a0 + n);
            if (n) check(ibk_io_read_anchor_file(a.c_str(), nv, offset, anchor_idx.data() + a0, n, &n));
            offset += nv;
        }
        num_vertices = offset;
    }

    // positions, force elements and target positions into a device-resident method object
    template <class Method>
    void registerWith(Method& ib) const
    {
        ib.setPositions(X);
        if (!spring_master.empty()) ib.registerSprings(spring_master, spring_slave, spring_kappa, spring_rest);
        if (!beam_curr.empty()) ib.registerBeams(beam_curr, beam_next, beam_prev, beam_rigidity, beam_curvature);
        if (!target_idx.empty())
        {
            std::vector<double> X0(target_idx.size() * d_ndim); // IBTargetPointForceSpec: the initial position
            for (size_t k = 0; k < target_idx.size(); ++k)
                for (int d = 0; d < d_ndim; ++d) X0[k * d_ndim + d] = X[(size_t)target_idx[k] * d_ndim + d];
            ib.registerTargetPoints(target_idx, target_kappa, target_eta, X0);
        }
    }

    int num_vertices = 0;
    std::vector<double> X; // [num_vertices][ndim]
    std::vector<int> spring_master, spring_slave, spring_fcn;
    std::vector<double> spring_kappa, spring_rest;
    std::vector<int> beam_prev, beam_curr, beam_next;
    std::vector<double> beam_rigidity, beam_curvature;
    std::vector<int> target_idx;
    std::vector<double> target_kappa, target_eta;
    std::vector<int> anchor_idx;

private:
    static void check(int rc)
    {
        if (rc != IBK_OK) throw std::runtime_error(std::string("IBStandardInitializerB200: ") + ibk_io_last_error());
    }
    static void grow(int n, std::vector<int>& a, std::vector<int>& b, std::vector<int>& c)
    {
        a.resize(a.size() + n);
        b.resize(b.size() + n);
        c.resize(c.size() + n);
    }
    int d_ndim;
};
} // namespace IBAMR_B200
