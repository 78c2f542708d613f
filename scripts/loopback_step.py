"""Two ranks as two contexts of ONE process on ONE GPU (loopback communicator of libibk.so): the N = 2 benchmark step
(one 512^3 patch + 2^23 markers per rank, 2 x 1 x 1, overlapped exchange).  For looking at the exchange kernels with ncu,
which cannot follow a torchrun job.  usage: loopback_step.py [cells] [log2_markers] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ibamr_b200 import api, halo  # noqa: E402
from bench import splitmix_unit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 23)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
world, pg = 2, (2, 1, 1)
patches = halo.cartesian_patches(3, pg, (n, n, n))
dom = tuple(n * pg[d] for d in range(3))
ctxs = [api.Context(0) for _ in range(world)]
halo.CommExchange.init_loopback(ctxs)
ibs, hxs = [], []
h = 1.0 / n
for r in range(world):
    me = patches[r]
    ib = api.IBMethodB200(3, (0, 0, 0), tuple(d - 1 for d in dom), (0.0,) * 3, tuple(float(p) for p in pg), (1, 1, 1),
                          [(me.lower, me.upper)], kernel_fcn="IB_4", ctx=ctxs[r])
    idx = np.arange(N, dtype=np.uint64) + np.uint64(r) * np.uint64(N)
    X = np.stack([(me.lower[d] + n * splitmix_unit(7 + d, idx)) * h for d in range(3)], axis=1)
    F = np.stack([2.0 * splitmix_unit(1 + d, idx) - 1.0 for d in range(3)], axis=1)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    ibs.append(ib)
    hxs.append(halo.CommExchange(ib, patches))


def step_pipelined():
    for r in range(world):
        hxs[r].fill_post()
    for r in range(world):
        c = ctxs[r]
        c.check(c.lib.ibk_spread_begin(c.h))
        ibs[r].spreadForcePart(0)
        hxs[r].accumulate_post()
        ibs[r].halo("f")
    for r in range(world):
        ibs[r].halo("u")
        hxs[r].fill_finish()
        ibs[r].interpolateVelocityPart(0)
    for r in range(world):
        hxs[r].accumulate_finish()
        ctxs[r].check(ctxs[r].lib.ibk_spread_end(ctxs[r].h))


def step_split():
    for r in range(world):
        c = ctxs[r]
        c.check(c.lib.ibk_spread_begin(c.h))
        ibs[r].spreadForcePart(2)
        hxs[r].accumulate_post()
        ibs[r].spreadForcePart(1)
        ibs[r].halo("f")
    for r in range(world):
        hxs[r].accumulate_finish()
        ctxs[r].check(ctxs[r].lib.ibk_spread_end(ctxs[r].h))
    for r in range(world):
        hxs[r].fill_post()
        ibs[r].halo("u")
        ibs[r].interpolateVelocityPart(1)
    for r in range(world):
        hxs[r].fill_finish()
        ibs[r].interpolateVelocityPart(2)


step = step_split if os.environ.get("IBK_BENCH_SEQUENCE") == "split" else step_pipelined
for _ in range(2):
    step()
for c in ctxs:
    c.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
for c in ctxs:
    c.synchronize()
print(f"loopback: {world} ranks on one GPU, {(time.perf_counter() - t0) / steps * 1e3:.3f} ms per step (both ranks, serialised on the device)")
