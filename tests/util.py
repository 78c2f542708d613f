"""Shared helpers for the tests: the reference tests' RNG stream and golden-file parsers."""
import re

import numpy as np


def std_uniform_stream(seed, count, lo=0.0, hi=1.0):
    """`count` draws of std::uniform_real_distribution<double>(lo, hi) over std::mt19937(seed)
    as libstdc++ produces them (generate_canonical<double,53>: two 32-bit draws per double).
    This is the stream tests/interpolate/interpolate_01.cpp:162-177 and
    tests/IBTK/ghost_accumulation_01.cpp:131-135 use."""
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2**32, size=2 * count, dtype=np.uint64)
    g0 = raw[0::2].astype(np.float64)
    g1 = raw[1::2].astype(np.float64)
    c = (g0 + g1 * 4294967296.0) / 18446744073709551616.0
    return c * (hi - lo) + lo


def splitmix64_unit(seed, idx):
    """u(i) = (splitmix64(seed ^ i) >> 11) * 2^-53, the counter-based generator of SURVEY 8(d)."""
    x = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def read_interpolate_golden(path):
    rows = []
    with open(path) as f:
        next(f)
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(t) for t in line.split(",")])
    return np.array(rows)


_ARR = re.compile(r"array\(([-\d,]+)\)\s*=\s*(\S+)")


def read_ghost_accumulation_golden(path, ndim):
    """Returns [dict(x_lower, x_upper, comps={axis: {index tuple: value}})] per patch."""
    patches = []
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    i = 0
    cur = None
    axis = 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln.startswith("Rank:"):
            cur = dict(x_lower=None, x_upper=None, comps={})
            patches.append(cur)
            axis = 0
        elif ln.startswith("x lower:"):
            cur["x_lower"] = tuple(float(lines[i + 1 + d]) for d in range(ndim))
            i += ndim
        elif ln.startswith("x upper:"):
            cur["x_upper"] = tuple(float(lines[i + 1 + d]) for d in range(ndim))
            i += ndim
        elif ln.startswith("Array side normal"):
            axis = int(ln.split("=")[1])
        else:
            m = _ARR.match(ln)
            if m:
                idx = tuple(int(t) for t in m.group(1).split(","))
                cur["comps"].setdefault(axis, {})[idx] = float(m.group(2))
        i += 1
    return patches


def read_index_utilities_golden(path, ndim):
    """Returns [(point, level, box_lower, box_upper, index, contains)]."""
    out = []
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    i = 0
    point = None
    level = None
    box = None
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("Point ="):
            vals = [float(ln.split("=")[1])]
            for d in range(1, ndim):
                vals.append(float(lines[i + d]))
            point = tuple(vals)  # printed with 6 digits only; tests use the exact source values
            i += ndim - 1
        elif ln.strip().startswith("Level ="):
            level = int(ln.split("=")[1])
        elif ln.strip().startswith("Box"):
            nums = [int(t) for t in re.findall(r"-?\d+", ln.split("=")[1])]
            box = (tuple(nums[:ndim]), tuple(nums[ndim:]))
        elif ln.strip().startswith("Index"):
            idx = tuple(int(t) for t in re.findall(r"-?\d+", ln.split("=")[1]))
        elif ln.strip().startswith("contains"):
            out.append((point, level, box[0], box[1], idx, int(ln.split("=")[1])))
        i += 1
    return out
