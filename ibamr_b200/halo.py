"""Inter-process halo exchange for the patch-partitioned level (one process per GPU).

Replaces, across processes, the SAMRAI schedules the reference runs around the hot path:

  fill        u ghost cells <- owner interiors (copy), periodic wrap included
              u_ghost_fill_scheds[ln]->fillData, ibtk/src/lagrangian/LDataManager.cpp:744
  accumulate  owner interiors += every other copy of the DOF (ghost copies and the interior copy of a
              face shared by two patches), SAMRAIGhostDataAccumulator::accumulateGhostData reverse
              scatter, ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:327-334

Copies between patches of the SAME process are done by ibk_halo_local inside libibk.so.  What crosses a process
boundary is planned, packed, sent and unpacked inside libibk.so as well (csrc/ibk_comm.cu: ibk_halo_plan_*,
ibk_comm_*, ibk_halo_*_post / *_finish, ibk_migrate; NCCL over NVLink): CommExchange below only forwards the calls,
the way a C++ IBAMR rank would make them.  The plan is static (boxes do not change between regrids) and is derived
identically on both ends from the global box list, so no metadata is exchanged.  Unpack-adds run in a fixed order
(source rank, then item order), so sums are reproducible.

HaloPlan / HaloExchange expose the same C++ plan to the CPU tests: they run it over torch.distributed (gloo) with a
numpy stand-in for the pack / unpack kernels, so the planning and the message order are exercised without a GPU.
IbkBackend (device pack / unpack + torch.distributed messages) is the round-1 path, kept as a cross-check of the
library's own transport.
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
    """All inter-rank items of one level, for `fill` (u) and `accumulate` (f): a view of the C++ plan
    (ibk_halo_plan_create in libibk.so, host code: works without a GPU)."""

    def __init__(self, patches: list, domain_ncells, periodic, gcw, rank: int):
        import ctypes as C

        from . import _lib
        lib = _lib.load()
        self.patches, self.rank = patches, rank
        ndim = len(domain_ncells)
        self.ndim = ndim
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        lo, hi = i32([p.lower for p in patches]).reshape(-1), i32([p.upper for p in patches]).reshape(-1)
        rk = i32([p.rank for p in patches])
        h = C.c_void_p()
        rc = lib.ibk_halo_plan_create(ndim, len(patches), pi(lo), pi(hi), pi(rk), pi(i32(domain_ncells)), pi(i32([1 if x else 0 for x in periodic])),
                                      pi(i32(gcw)), int(rank), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ibk_halo_plan_create failed ({rc})")
        # local ids as the library numbers them: position among the patches of the same rank, in list order
        seen, by_local = {}, {}
        for p in patches:
            k = seen.get(p.rank, 0)
            seen[p.rank] = k + 1
            by_local[(p.rank, k)] = p
        self.fill, self.accum = {}, {}  # (src_rank, dst_rank) -> [Item]
        for t, table in ((0, self.fill), (1, self.accum)):
            for k in range(lib.ibk_halo_plan_messages(h, t)):
                src, dst, n, cnt = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
                lib.ibk_halo_plan_message(h, t, k, C.byref(src), C.byref(dst), C.byref(n), C.byref(cnt))
                n = n.value
                ax, sl, dl = (np.zeros(n, dtype=np.int32) for _ in range(3))
                slo, shi, dlo, dhi = (np.zeros(n * ndim, dtype=np.int32) for _ in range(4))
                lib.ibk_halo_plan_items(h, t, k, pi(ax), pi(sl), pi(dl), pi(slo), pi(shi), pi(dlo), pi(dhi))
                table[(src.value, dst.value)] = [
                    Item(int(ax[i]), by_local[(src.value, int(sl[i]))], by_local[(dst.value, int(dl[i]))],
                         tuple(int(v) for v in slo[i * ndim:(i + 1) * ndim]), tuple(int(v) for v in shi[i * ndim:(i + 1) * ndim]),
                         tuple(int(v) for v in dlo[i * ndim:(i + 1) * ndim]), tuple(int(v) for v in dhi[i * ndim:(i + 1) * ndim]))
                    for i in range(n)]
                assert sum(it.count for it in table[(src.value, dst.value)]) == cnt.value
        lib.ibk_halo_plan_destroy(h)

    def neighbours(self, table):
        send = sorted({dst for (src, dst) in table if src == self.rank})
        recv = sorted({src for (src, dst) in table if dst == self.rank})
        return send, recv

    def bytes_per_exchange(self, table):
        return 8 * sum(it.count for (src, dst), items in table.items() if src == self.rank for it in items)


class CommExchange:
    """The device path: the communicator of the context inside libibk.so (ibk_comm_*; NCCL, or loopback between contexts of
    one process) does the planning, packing, messaging and unpacking; this class only forwards the calls."""

    def __init__(self, ib, patches):
        import ctypes as C
        self.ib, self.C = ib, C
        ctx = ib.ctx
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        lo, hi = i32([p.lower for p in patches]).reshape(-1), i32([p.upper for p in patches]).reshape(-1)
        rk = i32([p.rank for p in patches])
        ctx.check(ctx.lib.ibk_comm_set_patches(ctx.h, len(patches), pi(lo), pi(hi), pi(rk)))

    @staticmethod
    def init_nccl(ctx, dist, torch, rank, world):
        """Rank 0 creates the NCCL id, torch.distributed (the host program's own messaging) hands it round."""
        import ctypes as C
        idbuf = (C.c_char * 128)()
        if rank == 0:
            ctx.check(ctx.lib.ibk_comm_unique_id(idbuf))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        ctx.check(ctx.lib.ibk_comm_init(ctx.h, C.c_char_p(raw), rank, world))

    @staticmethod
    def init_loopback(contexts):
        import ctypes as C
        arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
        contexts[0].check(contexts[0].lib.ibk_comm_init_loopback(arr, len(contexts)))

    def _call(self, name):
        ctx = self.ib.ctx
        ctx.check(getattr(ctx.lib, name)(ctx.h))

    def fill_post(self):
        self._call("ibk_halo_fill_post")

    def fill_finish(self):
        self._call("ibk_halo_fill_finish")

    def accumulate_post(self):
        self._call("ibk_halo_accumulate_post")

    def accumulate_finish(self):
        self._call("ibk_halo_accumulate_finish")

    def bytes(self, which):
        return int(self.ib.ctx.lib.ibk_halo_bytes(self.ib.ctx.h, which))

    # the unsplit forms
    def fill(self):
        self.fill_post()
        self.fill_finish()

    accumulate_begin = accumulate_post  # pack BEFORE any local accumulation touches the values ...
    accumulate_end = accumulate_finish  # ... add AFTER it

    def migrate(self, id_bound):
        C = self.C
        ns, nr = C.c_int(), C.c_int()
        ctx = self.ib.ctx
        ctx.check(ctx.lib.ibk_migrate(ctx.h, C.c_uint(int(id_bound)), C.byref(ns), C.byref(nr)))
        return ns.value, nr.value


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
