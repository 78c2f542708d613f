// ibk_io.cu -- readers for the ASCII structure files of IBStandardInitializer (host code):
// <base>.vertex, .spring, .beam, .target, .anchor.  Same grammar and the same validity rules as the
// reference's readers (src/IB/IBStandardInitializer.cpp): text after '!', '#' or '%' is a comment
// (discard_comments, :65-87); line 1 holds the entry count; indices are checked against [0, n_vertices);
// negative stiffnesses / rest lengths / rigidities are errors; duplicated springs, beams and target points
// are skipped; optional trailing fields take the reference's defaults.  Indices are returned in the global
// Lagrangian numbering (vertex_offset added, :474-477).  Instead of TBOX_ERROR the functions return
// IBK_ERR_INVALID and leave the message in ibk_io_last_error().
#include <cstdio>
#include <fstream>
#include <limits>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/ibk.h"

namespace
{
thread_local std::string g_io_err;

int io_fail(const std::string& msg)
{
    g_io_err = msg;
    return IBK_ERR_INVALID;
}

std::string discard_comments(const std::string& in)
{
    std::string out = in;
    for (char c : { '!', '#', '%' })
    {
        const size_t p = out.find(c);
        if (p != std::string::npos) out.erase(p);
    }
    return out;
}

// line 1 = count; then `count` lines handed to `row` (1-based file line number for messages)
template <class RowFn>
int read_counted(const char* path, const char* what, bool required, int* count_out, RowFn row)
{
    *count_out = 0;
    if (!path) return io_fail("null path");
    std::ifstream file(path);
    if (!file.is_open())
    {
        if (required) return io_fail(std::string("Cannot find required ") + what + " file: " + path);
        return IBK_OK; // optional file: "does not exist: skipping read" (:514-518)
    }
    std::string line;
    if (!std::getline(file, line)) return io_fail(std::string("Premature end to input file encountered before line 1 of file ") + path);
    int count = -1;
    {
        std::istringstream ls(discard_comments(line));
        if (!(ls >> count) || count <= 0) return io_fail(std::string("Invalid entry in input file encountered on line 1 of file ") + path);
    }
    for (int k = 0; k < count; ++k)
    {
        if (!std::getline(file, line))
            return io_fail("Premature end to input file encountered before line " + std::to_string(k + 2) + " of file " + path);
        std::istringstream ls(discard_comments(line));
        std::string why;
        if (!row(ls, why))
            return io_fail("Invalid entry in input file encountered on line " + std::to_string(k + 2) + " of file " + path +
                           (why.empty() ? "" : ": " + why));
    }
    *count_out = count;
    return IBK_OK;
}

bool read_index(std::istringstream& ls, int n_vertices, int& v, std::string& why)
{
    if (!(ls >> v)) return false;
    if (v < 0 || v >= n_vertices)
    {
        why = "vertex index " + std::to_string(v) + " is out of range";
        return false;
    }
    return true;
}
} // namespace

extern "C" const char* ibk_io_last_error(void)
{
    return g_io_err.c_str();
}

extern "C" int ibk_io_read_vertex_file(const char* path, int ndim, double* X, int capacity, int* n_vertices)
{
    if (!n_vertices || (ndim != 2 && ndim != 3)) return io_fail("bad arguments");
    int k = 0, count = 0;
    int rc = read_counted(path, "vertex", true, &count, [&](std::istringstream& ls, std::string&) {
        for (int d = 0; d < ndim; ++d)
        {
            double v;
            if (!(ls >> v)) return false;
            if (X && k < capacity) X[(size_t)k * ndim + d] = v;
        }
        ++k;
        return true;
    });
    *n_vertices = count;
    return rc;
}

extern "C" int ibk_io_read_spring_file(const char* path, int n_vertices, int vertex_offset, int* master, int* slave, double* kappa,
                                       double* rest_length, int* force_fcn_idx, int capacity, int* n_springs)
{
    if (!n_springs) return io_fail("bad arguments");
    std::set<std::pair<int, int>> seen;
    int kept = 0, count = 0;
    int rc = read_counted(path, "spring", false, &count, [&](std::istringstream& ls, std::string& why) {
        int a, b, fcn = 0;
        double k, r;
        if (!read_index(ls, n_vertices, a, why) || !read_index(ls, n_vertices, b, why)) return false;
        if (!(ls >> k)) return false;
        if (k < 0.0)
        {
            why = "spring constant is negative";
            return false;
        }
        if (!(ls >> r)) return false;
        if (r < 0.0)
        {
            why = "spring resting length is negative";
            return false;
        }
        if (!(ls >> fcn)) fcn = 0; // default force function (:445-448)
        a += vertex_offset;
        b += vertex_offset;
        if (a > b) std::swap(a, b); // the edge belongs to its smaller index (:478-481)
        if (!seen.insert({ a, b }).second) return true; // duplicate connection: skipped (:482-499)
        if (kept < capacity)
        {
            if (master) master[kept] = a;
            if (slave) slave[kept] = b;
            if (kappa) kappa[kept] = k;
            if (rest_length) rest_length[kept] = r;
            if (force_fcn_idx) force_fcn_idx[kept] = fcn;
        }
        ++kept;
        return true;
    });
    *n_springs = kept;
    return rc;
}

extern "C" int ibk_io_read_beam_file(const char* path, int n_vertices, int vertex_offset, int ndim, int* prev, int* curr, int* next,
                                     double* rigidity, double* curvature, int capacity, int* n_beams)
{
    if (!n_beams || (ndim != 2 && ndim != 3)) return io_fail("bad arguments");
    std::set<std::tuple<int, int, int>> seen;
    int kept = 0, count = 0;
    int rc = read_counted(path, "beam", false, &count, [&](std::istringstream& ls, std::string& why) {
        int p, c, n;
        double bend, curv[3] = { 0.0, 0.0, 0.0 };
        if (!read_index(ls, n_vertices, p, why) || !read_index(ls, n_vertices, c, why) || !read_index(ls, n_vertices, n, why))
            return false;
        if (!(ls >> bend)) return false;
        if (bend < 0.0)
        {
            why = "beam constant is negative";
            return false;
        }
        bool found = false;
        for (int d = 0; d < ndim; ++d) // curvature: all NDIM components or none (:874-893)
        {
            double v;
            if (!(ls >> v))
            {
                if (found)
                {
                    why = "incomplete beam curvature specification";
                    return false;
                }
            }
            else
            {
                found = true;
                curv[d] = v;
            }
        }
        p += vertex_offset;
        c += vertex_offset;
        n += vertex_offset;
        if (!seen.insert(std::make_tuple(c, n, p)).second) return true; // duplicate: skipped (:950-969)
        if (kept < capacity)
        {
            if (prev) prev[kept] = p;
            if (curr) curr[kept] = c;
            if (next) next[kept] = n;
            if (rigidity) rigidity[kept] = bend;
            if (curvature)
                for (int d = 0; d < ndim; ++d) curvature[(size_t)kept * ndim + d] = curv[d];
        }
        ++kept;
        return true;
    });
    *n_beams = kept;
    return rc;
}

extern "C" int ibk_io_read_target_file(const char* path, int n_vertices, int vertex_offset, int* idx, double* kappa, double* eta,
                                       int capacity, int* n_targets)
{
    if (!n_targets) return io_fail("bad arguments");
    std::set<int> seen;
    int kept = 0, count = 0;
    int rc = read_counted(path, "target point", false, &count, [&](std::istringstream& ls, std::string& why) {
        int n;
        double k, e;
        if (!read_index(ls, n_vertices, n, why)) return false;
        if (!seen.insert(n).second) return true; // duplicate target point: skipped (:1412-1418)
        if (!(ls >> k)) return false;
        if (k < 0.0)
        {
            why = "target point spring constant is negative";
            return false;
        }
        if (!(ls >> e)) e = 0.0; // damping is optional (:1441-1444)
        if (e < 0.0)
        {
            why = "target point damping coefficient is negative";
            return false;
        }
        if (kept < capacity)
        {
            if (idx) idx[kept] = n + vertex_offset;
            if (kappa) kappa[kept] = k;
            if (eta) eta[kept] = e;
        }
        ++kept;
        return true;
    });
    *n_targets = kept;
    return rc;
}

extern "C" int ibk_io_read_anchor_file(const char* path, int n_vertices, int vertex_offset, int* idx, int capacity, int* n_anchors)
{
    if (!n_anchors) return io_fail("bad arguments");
    std::set<int> seen;
    int kept = 0, count = 0;
    int rc = read_counted(path, "anchor point", false, &count, [&](std::istringstream& ls, std::string& why) {
        int n;
        if (!read_index(ls, n_vertices, n, why)) return false;
        if (!seen.insert(n).second) return true; // duplicate anchor point: skipped
        if (kept < capacity && idx) idx[kept] = n + vertex_offset;
        ++kept;
        return true;
    });
    *n_anchors = kept;
    return rc;
}
