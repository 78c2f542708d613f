// samrai_standins.h -- minimal stand-ins for the SAMRAI types that appear in the signatures of
// IBTK::LEInteractor and IBAMR::IBStrategy (SAMRAI is a third-party dependency that is not in this
// image).  They carry exactly what the hot path reads: boxes, ghost widths, patch geometry and raw
// array pointers in SAMRAI's ArrayData layout (Fortran order, ghosts included).
// In a real IBAMR build these are replaced by the SAMRAI headers; see INTEGRATION.md.
#pragma once
#include <algorithm>
#include <array>
#include <cstdio>
#include <map>
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
// tbox::Database: the keyed store restart data goes through (LData::putToDatabase / LData(Pointer<Database>),
// ibtk/src/lagrangian/LData.cpp:99-130, 186-209).  The stand-in keeps the four value kinds LData uses and can write itself
// to / read itself from a flat binary file (SAMRAI's HDF5 restart files are a third-party format).
struct Database
{
    std::map<std::string, std::string> strings;
    std::map<std::string, int> integers;
    std::map<std::string, std::vector<int>> integer_arrays;
    std::map<std::string, std::vector<double>> double_arrays;
    void putString(const std::string& k, const std::string& v)
    {
        strings[k] = v;
    }
    void putInteger(const std::string& k, int v)
    {
        integers[k] = v;
    }
    void putIntegerArray(const std::string& k, const int* v, int n)
    {
        integer_arrays[k].assign(v, v + n);
    }
    void putDoubleArray(const std::string& k, const double* v, int n)
    {
        double_arrays[k].assign(v, v + n);
    }
    std::string getString(const std::string& k) const
    {
        return strings.at(k);
    }
    int getInteger(const std::string& k) const
    {
        return integers.at(k);
    }
    void getIntegerArray(const std::string& k, int* v, int n) const
    {
        const auto& a = integer_arrays.at(k);
        if ((int)a.size() != n) throw std::runtime_error("Database: size of " + k);
        std::copy(a.begin(), a.end(), v);
    }
    void getDoubleArray(const std::string& k, double* v, int n) const
    {
        const auto& a = double_arrays.at(k);
        if ((int)a.size() != n) throw std::runtime_error("Database: size of " + k);
        std::copy(a.begin(), a.end(), v);
    }
    void writeToFile(const std::string& path) const
    {
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("Database: cannot write " + path);
        auto wstr = [&](const std::string& t) {
            const int n = (int)t.size();
            std::fwrite(&n, 4, 1, f);
            std::fwrite(t.data(), 1, t.size(), f);
        };
        int n = (int)strings.size();
        std::fwrite(&n, 4, 1, f);
        for (auto& kv : strings)
        {
            wstr(kv.first);
            wstr(kv.second);
        }
        n = (int)integers.size();
        std::fwrite(&n, 4, 1, f);
        for (auto& kv : integers)
        {
            wstr(kv.first);
            std::fwrite(&kv.second, 4, 1, f);
        }
        n = (int)integer_arrays.size();
        std::fwrite(&n, 4, 1, f);
        for (auto& kv : integer_arrays)
        {
            wstr(kv.first);
            const int m = (int)kv.second.size();
            std::fwrite(&m, 4, 1, f);
            std::fwrite(kv.second.data(), 4, kv.second.size(), f);
        }
        n = (int)double_arrays.size();
        std::fwrite(&n, 4, 1, f);
        for (auto& kv : double_arrays)
        {
            wstr(kv.first);
            const int m = (int)kv.second.size();
            std::fwrite(&m, 4, 1, f);
            std::fwrite(kv.second.data(), 8, kv.second.size(), f);
        }
        std::fclose(f);
    }
    static std::shared_ptr<Database> readFromFile(const std::string& path)
    {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("Database: cannot read " + path);
        auto db = std::make_shared<Database>();
        auto rint = [&]() {
            int v = 0;
            if (std::fread(&v, 4, 1, f) != 1) throw std::runtime_error("Database: truncated " + path);
            return v;
        };
        auto rstr = [&]() {
            std::string t((size_t)rint(), '\0');
            if (!t.empty() && std::fread(&t[0], 1, t.size(), f) != t.size()) throw std::runtime_error("Database: truncated " + path);
            return t;
        };
        for (int n = rint(); n > 0; --n)
        {
            const std::string k = rstr();
            db->strings[k] = rstr();
        }
        for (int n = rint(); n > 0; --n)
        {
            const std::string k = rstr();
            db->integers[k] = rint();
        }
        for (int n = rint(); n > 0; --n)
        {
            const std::string k = rstr();
            std::vector<int> a((size_t)rint());
            if (!a.empty() && std::fread(a.data(), 4, a.size(), f) != a.size()) throw std::runtime_error("Database: truncated " + path);
            db->integer_arrays[k] = a;
        }
        for (int n = rint(); n > 0; --n)
        {
            const std::string k = rstr();
            std::vector<double> a((size_t)rint());
            if (!a.empty() && std::fread(a.data(), 8, a.size(), f) != a.size()) throw std::runtime_error("Database: truncated " + path);
            db->double_arrays[k] = a;
        }
        std::fclose(f);
        return db;
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
