"""Multi-GPU parity worker (run under torchrun, one rank per GPU): patch-partitioned periodic level,
markers owned by the rank whose patch holds their cell, spreadForce / interpolateVelocity with the
NCCL halo exchange of libibk.so (ibk_comm_*), gathered and compared on rank 0 with the oracle's model of the reference path
(redundant ghost-region spreading, interiors kept; interpolation after a ghost fill).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py IB_4 [migrate]

With `migrate` the markers start on arbitrary ranks and are moved to their owners first
(ibk_migrate over the NCCL communicator of libibk.so; LDataManager.cpp:1824-1837).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ibamr_b200 import api, halo  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.util import splitmix64_unit  # noqa: E402


def periodic_side_field(pg, axis, ncell, seed):
    c = pg.side_coords(axis)
    f = np.sin(2 * np.pi * c[axis] / (ncell[axis] * pg.dx[axis])) * np.cos(2 * np.pi * c[(axis + 1) % 3] / (ncell[(axis + 1) % 3] * pg.dx[0]))
    gi = []
    for d in range(3):
        cnt = pg.upper[d] - pg.lower[d] + 1 + (1 if d == axis else 0) + 2 * pg.gcw[d]
        gi.append(np.mod(np.arange(cnt) + pg.lower[d] - pg.gcw[d], ncell[d]))
    mesh = np.meshgrid(*reversed(gi), indexing="ij")[::-1]
    lin = mesh[0] + ncell[0] * (mesh[1] + ncell[1] * mesh[2])
    return np.ascontiguousarray(f + 1e-3 * splitmix64_unit(seed + axis, lin.reshape(-1)).reshape(f.shape))


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "IB_4"
    migrate = "migrate" in sys.argv[2:]
    overlap = "overlap" in sys.argv[2:]
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 64 if overlap else 32  # with 64 cells per rank there are interior tiles
    pgrid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    patches = halo.cartesian_patches(3, pgrid, (n, n, n))
    dom = tuple(n * pgrid[d] for d in range(3))
    me = patches[rank]
    g = orc.min_ghost_width(kernel)
    xup = tuple(float(p) for p in pgrid)
    level = orc.Level(3, (0,) * 3, dom, (0.0,) * 3, xup, (1, 1, 1), [(p.lower, p.upper) for p in patches], (g,) * 3)
    N = 40000
    X = np.stack([xup[d] * splitmix64_unit(61 + d, np.arange(N)) for d in range(3)], axis=1)
    F = np.stack([2 * splitmix64_unit(71 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    ref = orc.bin_level(level, X)
    mine = np.nonzero(ref["owner"] == rank)[0]

    ib = api.IBMethodB200(3, (0,) * 3, tuple(d - 1 for d in dom), (0.0,) * 3, xup, (1, 1, 1), [(me.lower, me.upper)], gcw=g,
                          kernel_fcn=kernel, ctx=ctx)
    # the library's own multi-rank layer (csrc/ibk_comm.cu): NCCL communicator of the context, plan, pack, messages, unpack
    halo.CommExchange.init_nccl(ctx, dist, torch, rank, world)
    hx = halo.CommExchange(ib, patches)
    pg = level.patch_geom(rank)
    u = [periodic_side_field(pg, a, dom, 300) for a in range(3)]
    for a in range(3):
        garbage = u[a].copy()
        interior = tuple(slice(g, s - g) for s in garbage.shape)
        mask = np.ones(garbage.shape, bool)
        mask[interior] = False
        garbage[mask] = 1e30
        ib.grid_upload("u", 0, a, garbage)
        ib.grid_upload("f", 0, a, np.full(pg.side_shape(a), 0.25))
    if migrate:
        # start from an arbitrary distribution (index mod world) and let the markers find their owners
        start = np.arange(rank, N, world)
        ib.setPositions(X[start])
        ib.setLData("F", F[start])
        ib.setIds(start, N)
        ib.beginDataRedistribution()
        n_sent, n_recv = hx.migrate(N)
        ib.n_markers = int(ctx.lib.ibk_markers_count(ctx.h))
        ib.beginDataRedistribution()
        assert n_sent > 0 and n_recv > 0
        assert np.array_equal(ib.getIds(), mine), "after the migration a rank holds exactly the markers of its patch"
        assert np.array_equal(ib.getLData("X"), X[mine]) and np.array_equal(ib.getLData("F"), F[mine])
        assert hx.migrate(N) == (0, 0)
        ib.beginDataRedistribution()
    else:
        ib.setPositions(X[mine])
        ib.setLData("F", F[mine])
        ib.beginDataRedistribution()
    if "pipelined" in sys.argv[2:]:
        # the sequence bench.py times at N > 1: u ghosts travel during the spread, f ghost contributions during the interpolation
        ctx.check(ctx.lib.ibk_comm_set_reserved_sms(ctx.h, 0))
        hx.fill_post()
        ctx.check(ctx.lib.ibk_spread_begin(ctx.h))
        ib.spreadForcePart(0)
        hx.accumulate_post()
        ib.halo("f")
        ib.halo("u")
        hx.fill_finish()
        ib.interpolateVelocityPart(0)
        hx.accumulate_finish()
        ctx.check(ctx.lib.ibk_spread_end(ctx.h))
    elif overlap:
        # the exchange in flight while the interior tiles are processed (ibk_*_part, HaloExchange.*_post/_finish)
        ctx.check(ctx.lib.ibk_spread_begin(ctx.h))
        ib.spreadForcePart(2)
        hx.accumulate_post()
        ib.spreadForcePart(1)
        ib.halo("f")
        hx.accumulate_finish()
        ctx.check(ctx.lib.ibk_spread_end(ctx.h))
        hx.fill_post()
        ib.halo("u")
        ib.interpolateVelocityPart(1)
        hx.fill_finish()
        ib.interpolateVelocityPart(2)
    else:
        # spreadForce with the inter-rank exchange interleaved (see include/ibk.h, ibk_spread_begin)
        ctx.check(ctx.lib.ibk_spread_begin(ctx.h))
        ib.spreadForce(accumulate_halo=False)
        hx.accumulate_begin()
        ib.halo("f")
        hx.accumulate_end()
        ctx.check(ctx.lib.ibk_spread_end(ctx.h))
        # interpolateVelocity with the inter-rank ghost fill
        ib.halo("u")
        hx.fill()
        ib.interpolateVelocity(fill_halo=False)
    U = ib.getLData("U")
    f = [ib.grid_download("f", 0, a) for a in range(3)]

    # reference model on this rank's patch
    lst = ref["patches"][rank]
    ii = lst["all_idx"][lst["interior_mask"]]
    sh = lst["all_shift"].reshape(-1, 3)[lst["interior_mask"]]
    U_ref = orc.side_interp(kernel, pg, u, X, ii, sh.reshape(-1))
    f_ref = [np.zeros(pg.side_shape(a)) for a in range(3)]
    orc.side_spread(kernel, pg, f_ref, X, F, lst["all_idx"], lst["all_shift"])
    err_u = float(np.max(np.abs(U - U_ref[mine])) / np.max(np.abs(U_ref[mine])))
    err_f = 0.0
    for a in range(3):
        sl = tuple(slice(g, s - g) for s in f[a].shape)
        err_f = max(err_f, float(np.max(np.abs(f[a][sl] - 0.25 - f_ref[a][sl])) / np.max(np.abs(f_ref[a][sl]))))
    t = torch.tensor([err_u, err_f], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MGPU_PARITY kernel={kernel} world={world} migrate={int(migrate)} overlap={int(overlap)} pipelined={int('pipelined' in sys.argv[2:])} interp_rel_err={t[0].item():.3e} spread_rel_err={t[1].item():.3e} "
              f"fill_bytes={hx.bytes(0)} accum_bytes={hx.bytes(1)}", flush=True)
        assert t[0].item() <= 1e-12 and t[1].item() <= 1e-12
    ib.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
