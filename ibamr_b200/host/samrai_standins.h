// samrai_standins.h -- minimal stand-ins for the SAMRAI types that appear in the signatures of
// IBTK::LEInteractor and IBAMR::IBStrategy (SAMRAI is a third-party dependency that is not in this
// image).  They carry exactly what the hot path reads: boxes, ghost widths, patch geometry and raw
// array pointers in SAMRAI's ArrayData layout (Fortran order, ghosts included).
// In a real IBAMR build these are replaced by the SAMRAI headers; see INTEGRATION.md.
#pragma once
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef NDIM
#define NDIM 3
#endif

namespace SAMRAI_standin
{
template <class T>
using Pointer = std::shared_ptr<T>; // tbox::Pointer

struct IntVector
{
    std::array<int, NDIM> v{};
    IntVector() = default;
    explicit IntVector(int a)
    {
        v.fill(a);
    }
    int& operator()(int d)
    {
        return v[d];
    }
    int operator()(int d) const
    {
        return v[d];
    }
    int min() const
    {
        int m = v[0];
        for (int d = 1; d < NDIM; ++d) m = v[d] < m ? v[d] : m;
        return m;
    }
};
using Index = IntVector;

struct Box
{
    Index lo, hi;
    Box() = default;
    Box(const Index& l, const Index& u) : lo(l), hi(u)
    {
    }
    const Index& lower() const
    {
        return lo;
    }
    const Index& upper() const
    {
        return hi;
    }
    bool operator==(const Box& o) const
    {
        return lo.v == o.lo.v && hi.v == o.hi.v;
    }
};

// geom::CartesianPatchGeometry
struct CartesianPatchGeometry
{
    std::array<double, NDIM> x_lower{}, x_upper{}, dx{};
    bool touches_regular_bdry = false;
    const double* getXLower() const
    {
        return x_lower.data();
    }
    const double* getXUpper() const
    {
        return x_upper.data();
    }
    const double* getDx() const
    {
        return dx.data();
    }
};

// hier::Patch
struct Patch
{
    Box box;
    Pointer<CartesianPatchGeometry> geom;
    const Box& getBox() const
    {
        return box;
    }
    Pointer<CartesianPatchGeometry> getPatchGeometry() const
    {
        return geom;
    }
};

// pdat::SideData<NDIM,double>: one array per axis over toSideBox(box, axis) grown by the ghost width
struct SideData
{
    Box box;
    IntVector gcw;
    int depth = 1;
    std::array<std::vector<double>, NDIM> data;
    SideData(const Box& b, int depth_, const IntVector& g) : box(b), gcw(g), depth(depth_)
    {
        for (int axis = 0; axis < NDIM; ++axis)
        {
            size_t n = (size_t)depth;
            for (int d = 0; d < NDIM; ++d) n *= (size_t)(b.hi(d) - b.lo(d) + 1 + (d == axis ? 1 : 0) + 2 * g(d));
            data[axis].assign(n, 0.0);
        }
    }
    int getDepth() const
    {
        return depth;
    }
    const IntVector& getGhostCellWidth() const
    {
        return gcw;
    }
    const Box& getBox() const
    {
        return box;
    }
    double* getPointer(int axis)
    {
        return data[axis].data();
    }
    const double* getPointer(int axis) const
    {
        return data[axis].data();
    }
    void fillAll(double v)
    {
        for (auto& a : data) std::fill(a.begin(), a.end(), v);
    }
};

// pdat::CellData<NDIM,double>
struct CellData
{
    Box box;
    IntVector gcw;
    int depth = 1;
    std::vector<double> data;
    CellData(const Box& b, int depth_, const IntVector& g) : box(b), gcw(g), depth(depth_)
    {
        size_t n = (size_t)depth;
        for (int d = 0; d < NDIM; ++d) n *= (size_t)(b.hi(d) - b.lo(d) + 1 + 2 * g(d));
        data.assign(n, 0.0);
    }
    int getDepth() const
    {
        return depth;
    }
    const IntVector& getGhostCellWidth() const
    {
        return gcw;
    }
    const Box& getBox() const
    {
        return box;
    }
    double* getPointer()
    {
        return data.data();
    }
    const double* getPointer() const
    {
        return data.data();
    }
};
// xfer::RefineSchedule / xfer::CoarsenSchedule, IBTK::RobinPhysBdryPatchStrategy: they appear in the signatures of
// IBStrategy::interpolateVelocity / spreadForce and of LDataManager::interp / spread.  The work they stand for -- ghost
// fill of u, ghost accumulation of f, physical-boundary fold-back -- is done inside libibk.so (ibk_halo_local, the
// communicator of the context, ibk_level_set_physical_boundaries), so the mirrors accept and ignore them.
struct RefineSchedule
{
};
struct CoarsenSchedule
{
};
struct RobinPhysBdryPatchStrategy
{
};
} // namespace SAMRAI_standin
