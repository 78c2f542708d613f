"""world_size-3 gloo test (CPU) of ibamr_b200/halo.py::MarkerMigration: the count / row messaging and the
bookkeeping around it, with a numpy stand-in for libibk.so's ibk_migrate_plan/pack/unpack.  The model checked
against: after a migration every rank holds exactly the markers whose cell (IndexUtilities::getCellIndex,
via the oracle) lies in its patches, with their X, U, F rows intact, in ascending Lagrangian index
(LDataManager.cpp:1475-1476 ownership rule, :1824-1837 scatter)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ibamr_b200 import halo
from oracle import oracle as orc
from tests.util import splitmix64_unit


class NumpyMarkerBackend:
    """Test-only backend with the method names of halo.IbkBackend that MarkerMigration uses."""

    def __init__(self, X, U, F, ids, cells_of):
        self.X, self.U, self.F, self.ids, self.cells_of = X, U, F, ids, cells_of
        self.order = None

    def alloc(self, n):
        return torch.zeros(max(n, 1), dtype=torch.float64)

    def view(self, buf, off, count):
        return buf[off:off + count]

    def isend(self, buf, dst):
        return dist.P2POp(dist.isend, buf, dst)

    def irecv(self, buf, src):
        return dist.P2POp(dist.irecv, buf, src)

    def run(self, ops):
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def migrate_plan(self, lower, upper, ranks, world, rank):
        cells = self.cells_of(self.X)
        dest = np.full(len(self.X), world, dtype=np.int64)
        for lo, hi, r in reversed(list(zip(lower, upper, ranks))):  # first matching patch wins
            inside = np.all((cells >= np.array(lo)) & (cells <= np.array(hi)), axis=1)
            dest[inside] = r
        assert not np.any(dest == world)
        self.keep = np.nonzero(dest == rank)[0]
        leave = np.nonzero(dest != rank)[0]
        self.order = leave[np.argsort(dest[leave], kind="stable")]
        return np.bincount(dest[leave], minlength=world)[:world]

    def migrate_pack(self, buf):
        o = self.order
        rows = np.concatenate([self.X[o], self.U[o], self.F[o], self.ids[o][:, None].astype(np.float64)], axis=1)
        buf[:rows.size].copy_(torch.from_numpy(rows.reshape(-1)))

    def migrate_unpack(self, buf, n_recv, id_bound):
        nd = self.X.shape[1]
        rows = buf.numpy()[:n_recv * (3 * nd + 1)].reshape(n_recv, 3 * nd + 1)
        X = np.concatenate([self.X[self.keep], rows[:, :nd]])
        U = np.concatenate([self.U[self.keep], rows[:, nd:2 * nd]])
        F = np.concatenate([self.F[self.keep], rows[:, 2 * nd:3 * nd]])
        ids = np.concatenate([self.ids[self.keep], rows[:, 3 * nd].astype(np.int64)])
        o = np.argsort(ids)
        self.X, self.U, self.F, self.ids = X[o], U[o], F[o], ids[o]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ndim, n = 3, 8
        patches = halo.cartesian_patches(ndim, (world, 1, 1), (n, n, n))
        dom = (n * world, n, n)
        xup = (float(world), 1.0, 1.0)
        N = 5000
        ids_all = np.arange(N)
        X = np.stack([xup[d] * splitmix64_unit(11 + d, ids_all) for d in range(ndim)], axis=1)
        U = np.stack([splitmix64_unit(21 + d, ids_all) for d in range(ndim)], axis=1)
        F = np.stack([splitmix64_unit(31 + d, ids_all) for d in range(ndim)], axis=1)

        def cells_of(Xl):
            return orc.get_cell_index(Xl, (0.0,) * ndim, xup, tuple(1.0 / n for _ in range(ndim)), (0,) * ndim,
                                      tuple(d - 1 for d in dom))

        start = ids_all[rank::world]  # an arbitrary initial distribution
        be = NumpyMarkerBackend(X[start].copy(), U[start].copy(), F[start].copy(), start.copy(), cells_of)
        mig = halo.MarkerMigration(patches, rank, world, be, N)
        n_sent, n_recv = mig.migrate()
        cells = cells_of(X)
        mine = np.nonzero((cells[:, 0] >= patches[rank].lower[0]) & (cells[:, 0] <= patches[rank].upper[0]))[0]
        ok = (np.array_equal(be.ids, mine) and np.array_equal(be.X, X[mine]) and np.array_equal(be.U, U[mine])
              and np.array_equal(be.F, F[mine]))
        # a second migration has nothing to move
        s2, r2 = mig.migrate()
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, n_sent, n_recv, s2, r2, len(be.ids)))
        if rank == 0:
            results.put(gathered)
    finally:
        dist.destroy_process_group()


def test_marker_migration_three_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, results)) for r in range(3)]
    for p in procs:
        p.start()
    gathered = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g[0] for g in gathered), gathered
    assert sum(g[1] for g in gathered) == sum(g[2] for g in gathered) > 0  # everything sent was received
    assert all(g[3] == 0 and g[4] == 0 for g in gathered)
    assert sum(g[5] for g in gathered) == 5000
