"""Inter-process halo exchange for the patch-partitioned level (one process per GPU).

Replaces, across processes, the SAMRAI schedules the reference runs around the hot path:

  fill        u ghost cells <- owner interiors (copy), periodic wrap included
              u_ghost_fill_scheds[ln]->fillData, ibtk/src/lagrangian/LDataManager.cpp:744
  accumulate  owner interiors += every other copy of the DOF (ghost copies and the interior copy of a
              face shared by two patches), SAMRAIGhostDataAccumulator::accumulateGhostData reverse
              scatter, ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:327-334

Copies between patches of the SAME process are done by ibk_halo_local inside libibk.so; this module
only plans and moves what crosses a process boundary: device pack kernel -> one message per
neighbour rank (torch.distributed batch_isend_irecv: NCCL over NVLink on the GPU box, gloo in the CPU
tests) -> device unpack(-add) kernel.  The plan is static (boxes do not change between regrids) and
is derived identically on both ends from the global box list, so no metadata is exchanged.
Unpack-adds run in a fixed order (source rank, then item order), so sums are reproducible.

The pack/unpack backend is injected: IbkBackend drives libibk.so on device buffers; the CPU tests
inject a numpy stand-in to exercise exactly this planning/messaging logic without a GPU.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class GlobalPatch:
    lower: tuple
    upper: tuple
    rank: int
    local_id: int  # index among the patches of `rank`


def _side_boxes(gp: GlobalPatch, axis, gcw):
    ndim = len(gp.lower)
    ilo = tuple(gp.lower)
    ihi = tuple(gp.upper[d] + (1 if d == axis else 0) for d in range(ndim))
    alo = tuple(ilo[d] - gcw[d] for d in range(ndim))
    ahi = tuple(ihi[d] + gcw[d] for d in range(ndim))
    return ilo, ihi, alo, ahi


def _intersect(lo1, hi1, lo2, hi2):
    lo = tuple(max(a, b) for a, b in zip(lo1, lo2))
    hi = tuple(min(a, b) for a, b in zip(hi1, hi2))
    if any(h < l for l, h in zip(lo, hi)):
        return None
    return lo, hi


def _box_minus(lo, hi, ilo, ihi):
    """Box difference [lo,hi] \\ [ilo,ihi] as disjoint boxes (highest dimension first)."""
    out = []
    ndim = len(lo)
    cur_lo, cur_hi = list(lo), list(hi)
    inter = _intersect(lo, hi, ilo, ihi)
    if inter is None:
        return [(tuple(lo), tuple(hi))]
    for d in reversed(range(ndim)):
        if cur_lo[d] < inter[0][d]:
            l, h = list(cur_lo), list(cur_hi)
            h[d] = inter[0][d] - 1
            out.append((tuple(l), tuple(h)))
        if cur_hi[d] > inter[1][d]:
            l, h = list(cur_lo), list(cur_hi)
            l[d] = inter[1][d] + 1
            out.append((tuple(l), tuple(h)))
        cur_lo[d], cur_hi[d] = inter[0][d], inter[1][d]
    return out


@dataclass
class Item:
    """One region moved between two patches: pack from (src patch, src box), unpack at (dst patch, dst box)."""
    axis: int
    src: GlobalPatch
    dst: GlobalPatch
    src_lo: tuple
    src_hi: tuple
    dst_lo: tuple
    dst_hi: tuple

    @property
    def count(self):
        return int(np.prod([h - l + 1 for l, h in zip(self.dst_lo, self.dst_hi)]))


class HaloPlan:
    """All inter-rank items of one level, for `fill` (u) and `accumulate` (f)."""

    def __init__(self, patches: list, domain_ncells, periodic, gcw, rank: int):
        self.patches, self.rank = patches, rank
        ndim = len(domain_ncells)
        self.ndim = ndim
        offs = [(-1, 0, 1) if periodic[d] else (0,) for d in range(ndim)]
        self.fill: dict = {}  # (src_rank, dst_rank) -> [Item]
        self.accum: dict = {}
        for axis in range(ndim):
            for dst in patches:
                d_ilo, d_ihi, d_alo, d_ahi = _side_boxes(dst, axis, gcw)
                for src in patches:
                    if src.rank == dst.rank:
                        continue  # same process: ibk_halo_local
                    if src.rank != rank and dst.rank != rank:
                        continue
                    s_ilo, s_ihi, s_alo, s_ahi = _side_boxes(src, axis, gcw)
                    for o in itertools.product(*offs):
                        sh = tuple(o[d] * domain_ncells[d] for d in range(ndim))
                        # ---- fill: ghost region of dst  <-  interior of src (shifted by sh)
                        si = (tuple(s_ilo[d] + sh[d] for d in range(ndim)), tuple(s_ihi[d] + sh[d] for d in range(ndim)))
                        for glo, ghi in _box_minus(d_alo, d_ahi, d_ilo, d_ihi):
                            r = _intersect(glo, ghi, *si)
                            if r:
                                self.fill.setdefault((src.rank, dst.rank), []).append(
                                    Item(axis, src, dst, tuple(r[0][d] - sh[d] for d in range(ndim)),
                                         tuple(r[1][d] - sh[d] for d in range(ndim)), r[0], r[1]))
                        # ---- accumulate: interior of dst  +=  every copy src holds of it (ghosts, shared face)
                        sa = (tuple(s_alo[d] + sh[d] for d in range(ndim)), tuple(s_ahi[d] + sh[d] for d in range(ndim)))
                        r = _intersect(d_ilo, d_ihi, *sa)
                        if r:
                            self.accum.setdefault((src.rank, dst.rank), []).append(
                                Item(axis, src, dst, tuple(r[0][d] - sh[d] for d in range(ndim)),
                                     tuple(r[1][d] - sh[d] for d in range(ndim)), r[0], r[1]))
        # first-match semantics of the fill (a ghost cell covered by two source interiors, i.e. a shared
        # face, carries the same value in both) and a canonical order everywhere
        for table in (self.fill, self.accum):
            for key in table:
                table[key].sort(key=lambda it: (it.axis, it.dst.local_id, it.src.rank, it.src.local_id, it.dst_lo, it.dst_hi))

    def neighbours(self, table):
        send = sorted({dst for (src, dst) in table if src == self.rank})
        recv = sorted({src for (src, dst) in table if dst == self.rank})
        return send, recv

    def bytes_per_exchange(self, table):
        return 8 * sum(it.count for (src, dst), items in table.items() if src == self.rank for it in items)


class HaloExchange:
    """Executes a HaloPlan.  backend provides:
         alloc(n) -> buffer;  view(buffer, offset, count) -> buffer slice
         pack(which, local_patch, axis, lo, hi, buffer_slice)
         unpack(which, local_patch, axis, lo, hi, buffer_slice, mode)   mode 0 copy / 1 add
         isend(buffer, dst_rank) / irecv(buffer, src_rank) -> P2POp-like, run(ops)
    """

    def __init__(self, plan: HaloPlan, backend):
        self.plan, self.be = plan, backend
        self.buf = {}
        for name, table in (("fill", plan.fill), ("accum", plan.accum)):
            for (src, dst), items in table.items():
                n = sum(it.count for it in items)
                self.buf[(name, src, dst)] = backend.alloc(n)

    def _pack(self, name, table, which):
        for (src, dst), items in sorted(table.items()):
            if src != self.plan.rank:
                continue
            off = 0
            buf = self.buf[(name, src, dst)]
            if hasattr(self.be, "pack_many"):  # one library call per message
                self.be.pack_many(which, self._batch(name, src, dst, items, True), buf)
                continue
            for it in items:
                self.be.pack(which, it.src.local_id, it.axis, it.src_lo, it.src_hi, self.be.view(buf, off, it.count))
                off += it.count

    def _batch(self, name, src, dst, items, sending):
        """The item arrays of one message (built once): local patch, axis, boxes, offsets into the buffer."""
        key = (name, src, dst, sending)
        cache = self.__dict__.setdefault("_batches", {})
        if key not in cache:
            import numpy as np
            off, offs = 0, []
            for it in items:
                offs.append(off)
                off += it.count
            side = (lambda it: (it.src.local_id, it.src_lo, it.src_hi)) if sending else (lambda it: (it.dst.local_id, it.dst_lo, it.dst_hi))
            cache[key] = (np.array([side(it)[0] for it in items], dtype=np.int32), np.array([it.axis for it in items], dtype=np.int32),
                          np.ascontiguousarray([side(it)[1] for it in items], dtype=np.int32).reshape(-1),
                          np.ascontiguousarray([side(it)[2] for it in items], dtype=np.int32).reshape(-1),
                          np.array(offs, dtype=np.int64))
        return cache[key]

    def _ops(self, name, table):
        ops = []
        for (src, dst) in sorted(table):
            if src == self.plan.rank:
                ops.append(self.be.isend(self.buf[(name, src, dst)], dst))
            elif dst == self.plan.rank:
                ops.append(self.be.irecv(self.buf[(name, src, dst)], src))
        return ops

    def _exchange(self, name, table):
        self.be.run(self._ops(name, table))

    def _unpack(self, name, table, which, mode):
        for (src, dst), items in sorted(table.items()):  # ascending source rank: fixed add order
            if dst != self.plan.rank:
                continue
            off = 0
            buf = self.buf[(name, src, dst)]
            if hasattr(self.be, "unpack_many"):
                self.be.unpack_many(which, self._batch(name, src, dst, items, False), buf, mode)
                continue
            for it in items:
                self.be.unpack(which, it.dst.local_id, it.axis, it.dst_lo, it.dst_hi, self.be.view(buf, off, it.count), mode)
                off += it.count

    # u: owner interiors -> remote ghosts
    def fill(self):
        self._pack("fill", self.plan.fill, 0)
        self._exchange("fill", self.plan.fill)
        self._unpack("fill", self.plan.fill, 0, 0)

    # f: pack BEFORE any local accumulation touches the values, add AFTER it (see module docstring)
    def accumulate_begin(self):
        self._pack("accum", self.plan.accum, 1)

    def accumulate_end(self):
        self._exchange("accum", self.plan.accum)
        self._unpack("accum", self.plan.accum, 1, 1)

    # The same two operations split so that the messages are in flight while the caller launches the tiles that
    # do not touch the exchanged regions (ibk_*_part): post = pack + start the messages, finish = wait + unpack.
    def fill_post(self):
        self._pack("fill", self.plan.fill, 0)
        self._pending_fill = self.be.post(self._ops("fill", self.plan.fill))

    def fill_finish(self):
        self.be.wait(self._pending_fill)
        self._unpack("fill", self.plan.fill, 0, 0)

    def accumulate_post(self):
        self._pack("accum", self.plan.accum, 1)
        self._pending_accum = self.be.post(self._ops("accum", self.plan.accum))

    def accumulate_finish(self):
        self.be.wait(self._pending_accum)
        self._unpack("accum", self.plan.accum, 1, 1)


class IbkBackend:
    """Device backend: torch CUDA buffers, libibk.so pack/unpack kernels, torch.distributed P2P."""

    def __init__(self, ib, dist, torch):
        import ctypes as C
        self.ib, self.dist, self.torch, self.C = ib, dist, torch, C
        self.which = {0: 0, 1: 1}

    def alloc(self, n):
        return self.torch.empty(max(n, 1), dtype=self.torch.float64, device="cuda")

    def view(self, buf, off, count):
        return buf[off:off + count]

    def _box(self, lo, hi):
        import numpy as np
        return np.ascontiguousarray(lo, dtype=np.int32), np.ascontiguousarray(hi, dtype=np.int32)

    def pack(self, which, patch, axis, lo, hi, buf):
        C = self.C
        l, h = self._box(lo, hi)
        ctx = self.ib.ctx
        ctx.check(ctx.lib.ibk_halo_pack(ctx.h, which, patch, axis, l.ctypes.data_as(C.POINTER(C.c_int)),
                                        h.ctypes.data_as(C.POINTER(C.c_int)), C.c_void_p(buf.data_ptr())))

    def unpack(self, which, patch, axis, lo, hi, buf, mode):
        C = self.C
        l, h = self._box(lo, hi)
        ctx = self.ib.ctx
        ctx.check(ctx.lib.ibk_halo_unpack(ctx.h, which, patch, axis, l.ctypes.data_as(C.POINTER(C.c_int)),
                                          h.ctypes.data_as(C.POINTER(C.c_int)), C.c_void_p(buf.data_ptr()), mode))

    def migrate_plan(self, lower, upper, ranks, world, rank):
        return self.ib.migrate_plan(lower, upper, ranks, world, rank)

    def migrate_pack(self, buf):
        self.ib.ctx.synchronize()  # the buffer was created on torch's stream
        self.ib.migrate_pack(buf.data_ptr())
        self.ib.ctx.synchronize()

    def migrate_unpack(self, buf, n_recv, id_bound):
        self.torch.cuda.synchronize()
        self.ib.migrate_unpack(buf.data_ptr(), n_recv, id_bound)
        self.ib.ctx.synchronize()

    def pack_many(self, which, batch, buf):
        C = self.C
        patch, axis, lo, hi, offs = batch
        ctx = self.ib.ctx
        pi = C.POINTER(C.c_int)
        ctx.check(ctx.lib.ibk_halo_pack_many(ctx.h, which, len(patch), patch.ctypes.data_as(pi), axis.ctypes.data_as(pi),
                                             lo.ctypes.data_as(pi), hi.ctypes.data_as(pi),
                                             offs.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_void_p(buf.data_ptr())))

    def unpack_many(self, which, batch, buf, mode):
        C = self.C
        patch, axis, lo, hi, offs = batch
        ctx = self.ib.ctx
        pi = C.POINTER(C.c_int)
        ctx.check(ctx.lib.ibk_halo_unpack_many(ctx.h, which, len(patch), patch.ctypes.data_as(pi), axis.ctypes.data_as(pi),
                                               lo.ctypes.data_as(pi), hi.ctypes.data_as(pi),
                                               offs.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_void_p(buf.data_ptr()), mode))

    def isend(self, buf, dst):
        return self.dist.P2POp(self.dist.isend, buf, dst)

    def irecv(self, buf, src):
        return self.dist.P2POp(self.dist.irecv, buf, src)

    def run(self, ops):
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()

    def post(self, ops):
        """Starts the messages (NCCL: on its own stream, after the work already queued on the current one)."""
        return self.dist.batch_isend_irecv(ops) if ops else []

    def wait(self, reqs):
        for r in reqs:
            r.wait()  # NCCL: the current stream waits, the host does not


class MarkerMigration:
    """Moves the markers whose cell now lies in another rank's patch to that rank: the scatter of
    LDataManager::endDataRedistribution (LDataManager.cpp:1519-1959, VecScatter :1824-1837; node counts are
    exchanged like computeNodeOffsets' allGather, :3058).  The backend plans / packs / unpacks (libibk.so's
    ibk_migrate_* on the device, a numpy stand-in in the CPU tests); this class does the messaging: one count
    message and one row message per peer.  Call after a re-bin; re-bin again afterwards."""

    def __init__(self, patches: list, rank: int, world: int, backend, id_bound: int):
        self.rank, self.world, self.backend, self.id_bound = rank, world, backend, int(id_bound)
        self.lower = [list(p.lower) for p in patches]
        self.upper = [list(p.upper) for p in patches]
        self.ranks = [p.rank for p in patches]
        self.ndim = len(patches[0].lower)
        self.width = 3 * self.ndim + 1

    def migrate(self):
        """Returns (markers sent, markers received)."""
        b = self.backend
        send_counts = [int(c) for c in b.migrate_plan(self.lower, self.upper, self.ranks, self.world, self.rank)]
        peers = [r for r in range(self.world) if r != self.rank]
        # counts
        cs = {r: b.alloc(1) for r in peers}
        cr = {r: b.alloc(1) for r in peers}
        for r in peers:
            cs[r].fill_(float(send_counts[r]))
        b.run([b.irecv(cr[r], r) for r in peers] + [b.isend(cs[r], r) for r in peers])
        recv_counts = [0] * self.world
        for r in peers:
            recv_counts[r] = int(round(float(cr[r][0].item())))
        n_send, n_recv = sum(send_counts), sum(recv_counts)
        # rows
        sbuf = b.alloc(n_send * self.width)
        rbuf = b.alloc(n_recv * self.width)
        b.migrate_pack(sbuf)
        ops, so, ro = [], 0, 0
        for r in range(self.world):
            if recv_counts[r]:
                ops.append(b.irecv(b.view(rbuf, ro * self.width, recv_counts[r] * self.width), r))
            ro += recv_counts[r]
        for r in range(self.world):
            if send_counts[r]:
                ops.append(b.isend(b.view(sbuf, so * self.width, send_counts[r] * self.width), r))
            so += send_counts[r]
        b.run(ops)
        b.migrate_unpack(rbuf, n_recv, self.id_bound)
        return n_send, n_recv


def cartesian_patches(ndim, ranks_per_dim, cells_per_rank):
    """The weak-scaling layout: one patch per rank on a px x py x pz process grid (x fastest)."""
    out = []
    dims = list(ranks_per_dim) + [1] * (3 - len(ranks_per_dim))
    r = 0
    for kz in range(dims[2]):
        for ky in range(dims[1]):
            for kx in range(dims[0]):
                k = (kx, ky, kz)[:ndim]
                lo = tuple(k[d] * cells_per_rank[d] for d in range(ndim))
                hi = tuple((k[d] + 1) * cells_per_rank[d] - 1 for d in range(ndim))
                out.append(GlobalPatch(lo, hi, r, 0))
                r += 1
    return out
