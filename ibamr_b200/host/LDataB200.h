// LDataB200.h -- host mirror of IBTK::LData (ibtk/include/ibtk/LData.h:51-368) for a marker column that lives on the
// device (seam B2 of SURVEY.md 8(b)): callers that are not accelerated -- force generators, the Silo writer, restart --
// keep calling getLocalFormVecArray() / restoreArrays(); the AoS host copy behind those calls is refreshed lazily:
//   * the device side is newer after a kernel wrote the column (markDeviceModified(), called by IBMethodB200);
//   * the host side is newer after somebody asked for a writable array; restoreArrays() pushes it back
//     (LData::restoreArrays is where the reference hands the arrays back to PETSc, LData-inl.h).
// Rows are in Lagrangian (host-row) order, depth values per row, exactly the layout LData's local form has
// (no ghost rows: the device design is owner-only).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ibk.h"
#include "samrai_standins.h"

namespace IBTK_B200
{
class LDataB200
{
public:
    // column: IBK_COL_X ... IBK_COL_AUX; depth is the level's NDIM
    LDataB200(std::string name, ibk_ctx* ctx, int column, int depth) : d_name(std::move(name)), d_ctx(ctx), d_column(column), d_depth(depth)
    {
    }
    // Restart (ibtk/src/lagrangian/LData.cpp:99-130): the column is filled from the database's "vals".  The level of `ctx`
    // must hold num_local_nodes markers already (positions are restored first: an LData("X") record through
    // IBMethodB200::setPositions, then ibk_rebin).  Rows are host rows (Lagrangian order): the reference writes the PETSc
    // local order, which is a product of LDataManager's own restart data, not of LData.
    LDataB200(const std::shared_ptr<SAMRAI_standin::Database>& db, ibk_ctx* ctx, int column)
        : d_name(db->getString("d_name")), d_ctx(ctx), d_column(column), d_depth(db->getInteger("d_depth"))
    {
        const int num_local_nodes = db->getInteger("num_local_nodes"), num_ghost_nodes = db->getInteger("num_ghost_nodes");
        if (num_ghost_nodes != 0) throw std::runtime_error("LDataB200 " + d_name + ": ghost nodes in the restart record (owner-only design)");
        if (num_local_nodes != ibk_markers_count(ctx)) throw std::runtime_error("LDataB200 " + d_name + ": node count differs from the level's");
        d_host.assign((size_t)num_local_nodes * d_depth, 0.0);
        if (num_local_nodes > 0) db->getDoubleArray("vals", d_host.data(), d_depth * num_local_nodes);
        d_host_newer = true;
        d_device_newer = false;
        restoreArrays();
    }
    // LData::putToDatabase (LData.cpp:186-209): same keys; no ghost nodes here
    void putToDatabase(const std::shared_ptr<SAMRAI_standin::Database>& db)
    {
        const int num_local_nodes = getLocalNodeCount();
        db->putString("d_name", d_name);
        db->putInteger("d_depth", d_depth);
        db->putInteger("num_local_nodes", num_local_nodes);
        db->putInteger("num_ghost_nodes", 0);
        const double* vals = static_cast<const LDataB200*>(this)->getLocalFormVecArray();
        if (num_local_nodes > 0) db->putDoubleArray("vals", vals, d_depth * num_local_nodes);
        restoreArrays();
    }
    const std::string& getName() const
    {
        return d_name;
    }
    int getDepth() const
    {
        return d_depth;
    }
    int column() const // the marker column behind this LData (IBK_COL_*)
    {
        return d_column;
    }
    int getLocalNodeCount() const
    {
        return ibk_markers_count(d_ctx);
    }
    int getGhostNodeCount() const
    {
        return 0;
    }
    // a kernel (spread / interpolate / force / step) wrote the device column
    void markDeviceModified()
    {
        d_device_newer = true;
    }
    // LData::getLocalFormVecArray(): writable [n][depth] view; the host copy is assumed modified until restoreArrays()
    double* getLocalFormVecArray()
    {
        pull();
        d_host_newer = true;
        return d_host.data();
    }
    // read-only access does not dirty the host copy
    const double* getLocalFormVecArray() const
    {
        const_cast<LDataB200*>(this)->pull();
        return d_host.data();
    }
    double* getGhostedLocalFormVecArray()
    {
        return getLocalFormVecArray();
    }
    // LData::restoreArrays(): hand the arrays back; a modified host copy goes to the device
    void restoreArrays()
    {
        if (!d_host_newer) return;
        if ((int)d_host.size() != getLocalNodeCount() * d_depth) throw std::runtime_error("LDataB200 " + d_name + ": size changed under a checked-out array");
        check(ibk_markers_upload(d_ctx, d_column, d_host.data()));
        d_host_newer = false;
        d_device_newer = false;
    }
    bool hostCopyIsCurrent() const
    {
        return !d_device_newer && (int)d_host.size() == ibk_markers_count(d_ctx) * d_depth;
    }

private:
    void pull()
    {
        const size_t want = (size_t)getLocalNodeCount() * d_depth;
        if (d_host_newer)
        {
            if (d_device_newer) throw std::runtime_error("LDataB200 " + d_name + ": both the host array and the device column were modified");
            return;
        }
        if (!d_device_newer && d_host.size() == want) return;
        d_host.assign(want, 0.0);
        if (want) check(ibk_markers_download(d_ctx, d_column, d_host.data()));
        d_device_newer = false;
    }
    void check(int rc) const
    {
        if (rc != IBK_OK) throw std::runtime_error("LDataB200 " + d_name + ": " + ibk_last_error(d_ctx));
    }
    std::string d_name;
    ibk_ctx* d_ctx;
    int d_column, d_depth;
    std::vector<double> d_host;
    bool d_host_newer = false, d_device_newer = true;
};
} // namespace IBTK_B200
