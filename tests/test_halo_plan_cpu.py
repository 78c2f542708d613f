"""The C++ halo plan of libibk.so (ibk_halo_plan_create: host code, no GPU) executed in ONE process with numpy arrays for every
rank of the benchmark's process grids (2x1x1, 2x2x1, 2x2x2) and checked against a brute-force global model:
  fill        every ghost copy of a DOF another rank owns = the owner's interior value        (LDataManager.cpp:744)
  accumulate  every interior copy = the sum of ALL copies of the DOF on all ranks, ghost copies,
              periodic images and the second interior copy of a shared face included         (SAMRAIGhostDataAccumulator.cpp:327-334)
A DOF is a side index modulo the periodic domain (the face N of the last patch is the face 0 of the first)."""
import numpy as np
import pytest

from ibamr_b200 import halo


def _run(pgrid, cells, gcw, periodic, check_accumulate=True):
    ndim = 3
    patches = halo.cartesian_patches(ndim, pgrid, cells)
    world = len(patches)
    dom = tuple(cells[d] * pgrid[d] for d in range(ndim))
    g = (gcw,) * ndim

    def shape(axis):
        return tuple(reversed([cells[d] + (1 if d == axis else 0) + 2 * g[d] for d in range(ndim)]))

    def global_index(p, axis):
        """per dimension: the level index of every array element of patch p"""
        return [np.arange(shape(axis)[ndim - 1 - d]) + p.lower[d] - g[d] for d in range(ndim)]

    def key(p, axis, local_images_distinct=False):
        gi = global_index(p, axis)
        k, mul, inside = 0, 1, np.ones(shape(axis), bool)
        mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
        for d in range(ndim):
            n = dom[d]
            if periodic[d] and (pgrid[d] > 1 or not local_images_distinct):
                k = k + np.mod(mesh[d], n) * mul
            elif periodic[d]:
                # one rank along this dimension: images and the second copy of the patch's own boundary face are the LOCAL
                # halo's business (ibk_halo_local), the inter-rank plan must not touch them: keep them distinct DOFs
                k = k + (mesh[d] + 64) * mul
            else:
                k = k + (mesh[d] + 64) * mul
                inside = inside & (mesh[d] >= 0) & (mesh[d] <= n - 1 + (1 if d == axis else 0))
            mul *= 4096
        return k, inside

    def interior_mask(p, axis):
        m = np.zeros(shape(axis), bool)
        m[tuple(slice(g[d], g[d] + cells[d] + (1 if d == axis else 0)) for d in reversed(range(ndim)))] = True
        return m

    def region(p, lo, hi):
        return tuple(slice(lo[d] - (p.lower[d] - g[d]), hi[d] - (p.lower[d] - g[d]) + 1) for d in reversed(range(ndim)))

    plans = [halo.HaloPlan(patches, dom, periodic, g, r) for r in range(world)]
    rng = np.random.default_rng(7)
    # ---- fill
    u = [[None] * ndim for _ in range(world)]
    for r, p in enumerate(patches):
        for a in range(ndim):
            k, _ = key(p, a)
            arr = np.full(shape(a), 1e30)
            m = interior_mask(p, a)
            arr[m] = np.sin(0.001 * k[m] + a)
            u[r][a] = arr
    for r in range(world):
        for (src, dst), items in plans[r].fill.items():
            if dst != r:
                continue
            for it in items:
                u[dst][it.axis][region(patches[dst], it.dst_lo, it.dst_hi)] = u[src][it.axis][region(patches[src], it.src_lo, it.src_hi)]
    checked = 0
    for r, p in enumerate(patches):
        for a in range(ndim):
            k, inside = key(p, a)
            m = ~interior_mask(p, a) & inside
            # ghost copies of DOFs that some OTHER rank owns (with one patch per rank and >= 2 ranks along a periodic dimension a
            # ghost cell never is an image of the patch's own interior; along a dimension with one rank it is: those are local)
            remote = np.ones(shape(a), bool)
            mesh = np.meshgrid(*reversed(global_index(p, a)), indexing="ij")[::-1]
            for d in range(ndim):
                if pgrid[d] == 1:
                    remote &= (mesh[d] >= p.lower[d]) & (mesh[d] <= p.upper[d] + (1 if d == a else 0))
            m &= remote
            assert np.max(np.abs(u[r][a][m] - np.sin(0.001 * k[m] + a))) < 1e-14
            checked += int(m.sum())
    assert checked > 0
    if not check_accumulate:
        return plans
    # ---- accumulate.  (Along a periodic dimension with ONE rank the inter-rank plan and the local halo, ibk_halo_local, share
    # the work -- a remote ghost value at the patch's upper face lands on the image of that face, the local face sync then
    # copies it over -- so the brute-force model below is only complete when every periodic dimension has two ranks.)
    f0 = [[None] * ndim for _ in range(world)]
    for r, p in enumerate(patches):
        for a in range(ndim):
            arr = rng.standard_normal(shape(a))
            _, inside = key(p, a, local_images_distinct=True)
            arr[~inside] = 0.0
            mesh = np.meshgrid(*reversed(global_index(p, a)), indexing="ij")[::-1]
            for d in range(ndim):
                if pgrid[d] == 1:
                    arr[(mesh[d] < p.lower[d]) | (mesh[d] > p.upper[d] + (1 if d == a else 0))] = 0.0
            f0[r][a] = arr
    sums = [dict() for _ in range(ndim)]
    for r, p in enumerate(patches):
        for a in range(ndim):
            k, _ = key(p, a, local_images_distinct=True)
            kk, vv = k.reshape(-1), f0[r][a].reshape(-1)
            order = np.argsort(kk, kind="stable")
            uk, start = np.unique(kk[order], return_index=True)
            s = np.add.reduceat(vv[order], start)
            for key_, val in zip(uk.tolist(), s.tolist()):
                sums[a][key_] = sums[a].get(key_, 0.0) + val
    f = [[a.copy() for a in fr] for fr in f0]
    for r in range(world):
        recv = sorted((src, items) for (src, dst), items in plans[r].accum.items() if dst == r)
        for src, items in recv:  # ascending source rank
            for it in items:
                f[r][it.axis][region(patches[r], it.dst_lo, it.dst_hi)] += f0[src][it.axis][region(patches[src], it.src_lo, it.src_hi)]
    for r, p in enumerate(patches):
        for a in range(ndim):
            k, _ = key(p, a, local_images_distinct=True)
            m = interior_mask(p, a)
            want = np.array([sums[a][x] for x in k[m].tolist()])
            assert np.max(np.abs(f[r][a][m] - want)) < 1e-12
    return plans


@pytest.mark.parametrize("pgrid", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
@pytest.mark.parametrize("gcw", [3, 4])
def test_plan_of_the_benchmark_process_grids(pgrid, gcw):
    # the benchmark's levels: periodic in every dimension
    plans = _run(pgrid, (10, 9, 8), gcw, (1, 1, 1), check_accumulate=pgrid == (2, 2, 2))
    # every rank talks to every other rank (faces, edges and corners of a 2 x 2 x 2 periodic grid are all remote)
    send, recv = plans[0].neighbours(plans[0].accum)
    assert send == recv == list(range(1, len(plans)))
    # the same grids, periodic only where there are two ranks: fill and accumulate against the brute-force model
    _run(pgrid, (10, 9, 8), gcw, tuple(int(n > 1) for n in pgrid))


def test_plan_non_periodic():
    _run((2, 2, 2), (8, 8, 8), 3, (0, 0, 0))
