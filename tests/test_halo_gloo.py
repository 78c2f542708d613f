"""world_size-2 gloo tests (CPU) of the inter-process halo plan and messaging in ibamr_b200/halo.py.

The product moves device buffers with libibk.so's pack/unpack kernels over NCCL; here the same
HaloPlan / HaloExchange code runs with a numpy stand-in for the pack/unpack backend, so the planning
(which regions go where, periodic images, shared faces, ordering) and the messaging are what is
tested.  Expected values come from a brute-force global model of the two operations:
  fill        every ghost copy of a DOF = the owner's interior value      (LDataManager.cpp:744)
  accumulate  every interior copy = sum of ALL copies of the DOF          (SAMRAIGhostDataAccumulator.cpp:327-334)
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ibamr_b200 import halo


class NumpyBackend:
    """Test-only backend: arrays[which][local_patch][axis] are numpy arrays incl. ghosts, C-ordered
    ([n2,] n1, n0); regions are index boxes in the level's index space."""

    def __init__(self, arrays, lowers, gcw):
        self.arrays, self.lowers, self.gcw = arrays, lowers, gcw

    def alloc(self, n):
        return torch.zeros(max(n, 1), dtype=torch.float64)

    def view(self, buf, off, count):
        return buf[off:off + count]

    def _slices(self, patch, lo, hi):
        ndim = len(lo)
        return tuple(slice(lo[d] - (self.lowers[patch][d] - self.gcw[d]), hi[d] - (self.lowers[patch][d] - self.gcw[d]) + 1)
                     for d in reversed(range(ndim)))

    def pack(self, which, patch, axis, lo, hi, buf):
        buf.copy_(torch.from_numpy(np.ascontiguousarray(self.arrays[which][patch][axis][self._slices(patch, lo, hi)]).reshape(-1)))

    def unpack(self, which, patch, axis, lo, hi, buf, mode):
        sl = self._slices(patch, lo, hi)
        a = self.arrays[which][patch][axis]
        v = buf.numpy().reshape(a[sl].shape)
        if mode == 0:
            a[sl] = v
        else:
            a[sl] += v

    def isend(self, buf, dst):
        return dist.P2POp(dist.isend, buf, dst)

    def irecv(self, buf, src):
        return dist.P2POp(dist.irecv, buf, src)

    def run(self, ops):
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def post(self, ops):
        return dist.batch_isend_irecv(ops) if ops else []

    def wait(self, reqs):
        for r in reqs:
            r.wait()


def _global_value(axis, idx, ncells, periodic):
    """A value that depends only on the DOF (global side index, periodic images identified)."""
    key = 0
    mul = 1
    for d, i in enumerate(idx):
        n = ncells[d]
        g = i % n if periodic[d] else i
        key += (g + 7) * mul
        mul *= 131
    return np.sin(0.37 * key + axis)


def _worker(rank, world, port, ndim, periodic, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = (6, 5, 4)[:ndim]
        gcw = (2,) * ndim
        grid = (2,) + (1,) * (ndim - 1)
        patches = halo.cartesian_patches(ndim, grid, cells)
        dom = tuple(cells[d] * grid[d] for d in range(ndim))
        me = patches[rank]
        plan = halo.HaloPlan(patches, dom, periodic, gcw, rank)

        def shape(axis):
            return tuple(reversed([cells[d] + (1 if d == axis else 0) + 2 * gcw[d] for d in range(ndim)]))

        def index_arrays(axis):
            return np.meshgrid(*[np.arange(shape(axis)[ndim - 1 - d]) + me.lower[d] - gcw[d] for d in reversed(range(ndim))],
                               indexing="ij")[::-1]

        # ---------------- fill: interiors hold the global function, ghosts garbage
        u = []
        for axis in range(ndim):
            idx = index_arrays(axis)
            val = _global_value(axis, idx, dom, periodic)
            a = np.full(shape(axis), 1e30)
            interior = tuple(slice(gcw[d], gcw[d] + cells[d] + (1 if d == axis else 0)) for d in reversed(range(ndim)))
            a[interior] = val[interior]
            u.append(a)
        f = []
        for axis in range(ndim):
            # accumulate: every copy carries a distinct contribution: rank-dependent function of the index
            idx = index_arrays(axis)
            f.append(np.cos(0.11 * sum((i + 3) * (k + 1) for k, i in enumerate(idx)) + rank + axis))
        be = NumpyBackend({0: [u], 1: [f]}, [me.lower], gcw)
        hx = halo.HaloExchange(plan, be)
        f_before = [a.copy() for a in f]
        if periodic[0] and ndim == 3:  # both exchanges in flight at once (the pipelined step of bench.py at N > 1)
            hx.fill_post()
            hx.accumulate_post()
            hx.fill_finish()
            hx.accumulate_finish()
        elif periodic[0]:  # the split (overlappable) forms
            hx.fill_post()
            hx.fill_finish()
            hx.accumulate_post()
            hx.accumulate_finish()
        else:
            hx.fill()
            hx.accumulate_begin()
            hx.accumulate_end()
        # gather everything on rank 0 for the brute-force check
        payload = dict(rank=rank, lower=me.lower, u=u, f=f, f_before=f_before)
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank == 0:
            ok = True
            msgs = []
            for g in gathered:
                r, lo = g["rank"], g["lower"]
                for axis in range(ndim):
                    ua = g["u"][axis]
                    it = np.ndindex(ua.shape)
                    for loc in it:
                        gi = tuple(loc[ndim - 1 - d] + lo[d] - gcw[d] for d in range(ndim))
                        # is this element inside the domain along the non-periodic dims, and is it owned remotely?
                        inside = all(periodic[d] or 0 <= gi[d] < dom[d] + (1 if d == axis else 0) for d in range(ndim))
                        if not inside:
                            continue
                        interior = all(lo[d] <= gi[d] <= lo[d] + cells[d] - 1 + (1 if d == axis else 0) for d in range(ndim))
                        # expected fill: ghost elements whose DOF lives in the OTHER rank's interior
                        other = gathered[1 - r]
                        olo = other["lower"]

                        def in_other_interior(shift):
                            return all(olo[d] <= gi[d] - shift[d] <= olo[d] + cells[d] - 1 + (1 if d == axis else 0)
                                       for d in range(ndim))

                        shifts = [tuple(s * dom[d] if d == 0 else 0 for d in range(ndim)) for s in (-1, 0, 1)] if periodic[0] else [
                            (0,) * ndim]
                        if not interior and any(in_other_interior(s) for s in shifts):
                            exp = _global_value(axis, gi, dom, periodic)
                            if abs(ua[loc] - exp) > 1e-14:
                                ok = False
                                msgs.append(("fill", r, axis, gi, ua[loc], exp))
                        # expected accumulate: interior element = own value + every copy the other rank holds
                        if interior:
                            exp = g["f_before"][axis][loc]
                            ob = other["f_before"][axis]
                            for s in shifts:
                                oi = tuple(gi[d] - s[d] - (olo[d] - gcw[d]) for d in range(ndim))
                                if all(0 <= oi[d] < ob.shape[ndim - 1 - d] for d in range(ndim)):
                                    exp += ob[tuple(reversed(oi))]
                            if abs(g["f"][axis][loc] - exp) > 1e-13:
                                ok = False
                                msgs.append(("accum", r, axis, gi, g["f"][axis][loc], exp))
            results.put((ok, msgs[:5], plan.bytes_per_exchange(plan.fill), plan.bytes_per_exchange(plan.accum)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("ndim,periodic", [(2, (1, 0)), (3, (1, 0, 0)), (2, (0, 0))])
def test_halo_exchange_two_ranks(ndim, periodic):
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ndim, periodic, results)) for r in range(2)]
    for p in procs:
        p.start()
    ok, msgs, fill_bytes, accum_bytes = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, msgs
    assert fill_bytes > 0 and accum_bytes > 0


def test_plan_symmetry_and_sizes():
    """Both ends derive the same items: what rank 0 plans to send to rank 1 is what rank 1 plans to receive."""
    patches = halo.cartesian_patches(3, (2, 2, 1), (8, 8, 8))
    dom = (16, 16, 8)
    plans = [halo.HaloPlan(patches, dom, (1, 1, 1), (3, 3, 3), r) for r in range(4)]
    for a in range(4):
        for b in range(4):
            if a == b:
                continue
            for name in ("fill", "accum"):
                ta, tb = getattr(plans[a], name), getattr(plans[b], name)
                ia, ib = ta.get((a, b), []), tb.get((a, b), [])
                assert [(i.axis, i.src_lo, i.src_hi, i.dst_lo, i.dst_hi) for i in ia] == \
                       [(i.axis, i.src_lo, i.src_hi, i.dst_lo, i.dst_hi) for i in ib]
    # face message of the fill: gcw * n1 * n2 per component and face (+ edges); sanity on the volume
    n = sum(i.count for i in plans[0].fill.get((1, 0), []))
    assert n >= 3 * 3 * 8 * 8
