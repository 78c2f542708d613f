#!/usr/bin/env python
"""bench.py -- IB_4 spread + interpolate throughput (markers/s) on B200, one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (weak scaling, BASELINE.json config 5's per-GPU shard): every GPU owns one 512^3 patch of a
periodic staggered grid and 2^23 markers uniformly distributed inside it; N GPUs form a
(2,1,1) / (2,2,1) / (2,2,2) process grid, i.e. 8 GPUs = 64M markers on 1024^3.  Kernel IB_4, fp64.
A step = one pass of the hot path on pre-binned markers:
    spreadForce         zero f ghosts, f += S[F] (owner-only, deterministic), halo accumulate
    interpolateVelocity halo fill of u, U = J[u]
`value` times the steps with inputs resident in HBM; `e2e` times the same step through the
C ABI with HOST buffers (X, F, u in; U, f out over PCIe, pinned memory) -- the drop-in situation where
the fluid solver stays on the CPU.  Inputs (2 x 3.3 GB of grid data per GPU) are far larger than the
126 MB L2, so no explicit L2 flush is needed between iterations.

--impl reference times the reference's own CPU path: the oracle's restatement of the Fortran kernels
driven the way the reference parallelises (one worker per patch with private arrays, redundant
ghost-region spreading), on all host cores (OpenMP threads; MPI is not in this image), on the GPU arm's N = 1
workload (512^3 cells, 2^23 markers).  The `cpu_baseline` object of the GPU arm's own line is a bounded
density-preserving sample of it (256^3 cells, 2^20 markers).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNEL = "IB_4"  # (--config C3 switches to IB_6: set in main())
METRIC = "IB_4 spread+interp markers/sec"
UNIT = "markers/s"


def splitmix_unit(seed, idx):
    x = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def process_grid(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# --------------------------------------------------------------------------------------------------
def cpu_baseline(steps=2, warmup=1, n=256, log2_markers=20):
    """The reference's CPU path (oracle port) on all host cores: uniform markers on an n^3 periodic grid, a few passes (bounded:
    a pass over the 512^3 / 2^23 workload takes about a second on 16 threads)."""
    from oracle import oracle as orc
    threads = orc.Baseline.threads()
    # one patch per worker, SAMRAI-style box decomposition of the n^3 sample
    np3 = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 16: (4, 2, 2), 32: (4, 4, 2), 64: (4, 4, 4)}
    t = 1
    while t * 2 <= threads and t * 2 in np3:
        t *= 2
    npatch = np3[t]
    N = 1 << log2_markers
    X = np.stack([splitmix_unit(7 + d, np.arange(N)) for d in range(3)], axis=1).copy()
    F = np.stack([2.0 * splitmix_unit(1 + d, np.arange(N)) - 1.0 for d in range(3)], axis=1).copy()
    U = np.zeros((N, 3))
    b = orc.Baseline(3, (n, n, n), npatch, orc.min_ghost_width(KERNEL), (0.0,) * 3, (1.0,) * 3, X, field_seed=0)
    for _ in range(warmup):
        b.step(KERNEL, F, U)
    t0 = time.perf_counter()
    for _ in range(steps):
        b.step(KERNEL, F, U)
    dt = (time.perf_counter() - t0) / steps
    b.close()
    return {"value": N / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n}^3 periodic staggered grid cut into {npatch[0]}x{npatch[1]}x{npatch[2]} patches (one worker each, "
                      f"gcw {orc.min_ghost_width(KERNEL)}, redundant ghost-region spreading), 2^{log2_markers} uniform markers, {KERNEL} spread+interp, "
                      f"{steps} timed passes; oracle/_ref unbuildable here (needs m4+gfortran+SAMRAI+PETSc+MPI)",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the configuration the GPU arm runs at N = 1 (one 512^3 patch group, 2^23 markers), on all host cores
    cb = cpu_baseline(steps=max(args.steps, 1), warmup=max(min(args.warmup, 2), 1), n=args.cells, log2_markers=args.log2_markers)
    cfg = workload_config(1, args)
    cfg["workload"] = (f"C5 shard on the host: one {args.cells}^3 periodic staggered grid + 2^{args.log2_markers} uniform markers, "
                       f"IB_4, fp64 -- the N = 1 workload of the GPU arm (the CPU arm does not scale with --gpus)")
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, args):
    pg = process_grid(n_gpus)
    n = args.cells
    name = {"C5": "C5 shard (weak scaling)", "C3": "C3 (IB_6, uniform + clustered shell; per-GPU shard at N > 1)",
            "C2": "C2 (spherical shell)"}[args.config]
    return {"workload": f"{name}: per GPU one {n}^3 patch of a periodic staggered grid + 2^{args.log2_markers} "
                        f"{args.markers} markers, {KERNEL}, fp64; process grid {pg[0]}x{pg[1]}x{pg[2]} "
                        f"(global {n * pg[0]}x{n * pg[1]}x{n * pg[2]}, {n_gpus * (1 << args.log2_markers)} markers)",
            "kernel": KERNEL, "cells_per_gpu": [n, n, n], "markers_per_gpu": 1 << args.log2_markers,
            "step": "spreadForce (ghost zero + spread + halo accumulate) + interpolateVelocity (halo fill + interp), markers pre-binned",
            "l2": "inputs larger than L2 (2 x 3.3 GB grid data per GPU vs 126 MB)"}


# --------------------------------------------------------------------------------------------------
# result checks (outside every timed region)
# --------------------------------------------------------------------------------------------------
def owned(arr, a, g, n):
    """The DOFs of component `a` this rank owns: interior cells, and along the axis the faces 0..n-1 (the upper face is the
    neighbour's / the periodic image of face 0).  arr is [z][y][x] with g ghost layers."""
    return arr[g:g + n, g:g + n, g:g + n]


def invariant_check(hf, hu, hF, hU, g, n, h, world, dist, torch):
    """Size-independent properties of the benchmark's own result: the spread conserves the total force,
    sum_owned f_a h^3 = sum_i F_ia, and spread and interpolation are discretely adjoint, <S F, u> h^3 = <F, J u>.
    hf = S[F] (f was zero), hU = J[u]: the host arrays the e2e step left behind.  Sums are all-reduced over the ranks."""
    vol = h ** 3
    parts = []
    for a in range(3):
        fa = owned(hf[a].numpy(), a, g, n)
        parts += [float(fa.sum(dtype=np.float64)) * vol, float(hF.numpy()[:, a].sum(dtype=np.float64)),
                  float(np.abs(hF.numpy()[:, a]).sum(dtype=np.float64))]
    sfu = sum(float(np.vdot(owned(hf[a].numpy(), a, g, n), owned(hu[a].numpy(), a, g, n))) for a in range(3)) * vol
    fju = float(np.vdot(hF.numpy(), hU.numpy()))
    scale = float(np.abs(hF.numpy() * hU.numpy()).sum(dtype=np.float64))
    v = np.array(parts + [sfu, fju, scale], dtype=np.float64)
    if world > 1:
        t = torch.from_numpy(v).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        v = t.cpu().numpy()
    force = [abs(v[3 * a] - v[3 * a + 1]) / v[3 * a + 2] for a in range(3)]
    adj = abs(v[9] - v[10]) / v[11]
    ok = bool(max(force) <= 1e-10 and adj <= 1e-10)
    return {"total_force_rel_err": [float(x) for x in force], "adjointness_rel_err": float(adj), "tolerance": 1e-10, "ok": ok,
            "what": "on the benchmark workload itself, all ranks: sum_owned f_a h^3 vs sum_i F_ia (relative to sum |F_ia|), "
                    "<S F, u> h^3 vs <F, J u> (relative to sum |F.U|)"}


def sample_parity(ctx_device, n=256, log2_markers=20):
    """The CPU arm's sample (n^3 periodic grid, 2^log2_markers uniform markers, IB_4) through the GPU path, compared value by
    value with the oracle's model of the reference (its patch-private arrays, redundant ghost-region spreading)."""
    from ibamr_b200 import api
    from oracle import oracle as orc
    threads = orc.Baseline.threads()
    t = 1
    while t * 2 <= min(threads, 8):
        t *= 2
    npatch = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[t]
    N = 1 << log2_markers
    X = np.stack([splitmix_unit(7 + d, np.arange(N)) for d in range(3)], axis=1).copy()
    F = np.stack([2.0 * splitmix_unit(1 + d, np.arange(N)) - 1.0 for d in range(3)], axis=1).copy()
    U_ref = np.zeros((N, 3))
    g = orc.min_ghost_width(KERNEL)
    b = orc.Baseline(3, (n, n, n), npatch, g, (0.0,) * 3, (1.0,) * 3, X, field_seed=0)
    b.zero_f()
    b.step(KERNEL, F, U_ref)
    # global interior arrays from the workers' patches
    w = [n // npatch[d] for d in range(3)]
    u_glob = [np.zeros(tuple(n + (1 if (2 - k) == a else 0) for k in range(3))) for a in range(3)]  # [z][y][x]
    f_glob = [np.zeros_like(u_glob[a]) for a in range(3)]
    q = 0
    for pz in range(npatch[2]):
        for py in range(npatch[1]):
            for px in range(npatch[0]):
                lo = (px * w[0], py * w[1], pz * w[2])
                for a in range(3):
                    shp = tuple(w[2 - k] + 2 * g + (1 if (2 - k) == a else 0) for k in range(3))
                    sl = tuple(slice(g, shp[k] - g) for k in range(3))
                    dst = tuple(slice(lo[2 - k], lo[2 - k] + w[2 - k] + (1 if (2 - k) == a else 0)) for k in range(3))
                    u_glob[a][dst] = b.patch_array("u", q, a).reshape(shp)[sl]
                    f_glob[a][dst] = b.patch_array("f", q, a).reshape(shp)[sl]
                q += 1
    b.close()
    ib = api.IBMethodB200(3, (0, 0, 0), (n - 1,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0, 0, 0), (n - 1,) * 3)],
                          kernel_fcn=KERNEL, ctx=api.Context(ctx_device))
    gg = ib.gcw[0]
    for a in range(3):
        arr = np.zeros(ib.side_shape(0, a))
        arr[tuple(slice(gg, s - gg) for s in arr.shape)] = u_glob[a]
        ib.grid_upload("u", 0, a, arr)
    ib.setPositions(X)
    ib.setLData("F", F)
    ib.beginDataRedistribution()
    ib.grid_fill("f", 0.0)
    ib.spreadForce(accumulate_halo=True)
    ib.interpolateVelocity(fill_halo=True)
    U = ib.getLData("U")
    err_f = 0.0
    for a in range(3):
        fa = ib.grid_download("f", 0, a)
        fa = fa[tuple(slice(gg, s - gg) for s in fa.shape)]
        err_f = max(err_f, float(np.max(np.abs(fa - f_glob[a])) / np.max(np.abs(f_glob[a]))))
    err_u = float(np.max(np.abs(U - U_ref)) / np.max(np.abs(U_ref)))
    ib.close()
    return {"max_rel_err_U": err_u, "max_rel_err_f": err_f, "tolerance": 1e-12, "ok": bool(err_u <= 1e-12 and err_f <= 1e-12),
            "what": f"{n}^3 periodic grid, 2^{log2_markers} uniform markers, {KERNEL}: GPU path (one resident patch, halo fill / "
                    f"accumulate) vs the oracle's reference model ({npatch[0]}x{npatch[1]}x{npatch[2]} patches), max-norm relative"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from ibamr_b200 import api, halo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    n, N = args.cells, 1 << args.log2_markers
    pg = process_grid(world)
    patches = halo.cartesian_patches(3, pg, (n, n, n))
    me = patches[rank]
    dom = tuple(n * pg[d] for d in range(3))
    ib = api.IBMethodB200(3, (0, 0, 0), tuple(d - 1 for d in dom), (0.0,) * 3, tuple(float(p) for p in pg), (1, 1, 1),
                          [(me.lower, me.upper)], kernel_fcn=KERNEL, ctx=ctx)
    g = ib.gcw[0]
    hx = None
    if world > 1:
        # the multi-rank layer of libibk.so: its own NCCL communicator (id handed round by torch.distributed), plan, pack,
        # messages on the context's communication stream, unpack
        halo.CommExchange.init_nccl(ctx, dist, torch, rank, world)
        hx = halo.CommExchange(ib, patches)

    # ---- synthetic inputs: uniform markers in the rank's patch, smooth + noisy velocity
    h = 1.0 / n
    idx = np.arange(N, dtype=np.uint64) + np.uint64(rank) * np.uint64(N)
    X = np.stack([(me.lower[d] + n * splitmix_unit(7 + d, idx)) * h for d in range(3)], axis=1)
    if args.markers == "shell":
        # diagnostic workload (BASELINE config C2's shape): a Fibonacci-lattice sphere of radius n/4 cells in the
        # middle of the rank's patch: a dense surface (tens of markers per cell) instead of uniform markers
        k = np.arange(N, dtype=np.float64) + 0.5
        phi = np.arccos(1.0 - 2.0 * k / N)
        th = np.pi * (1.0 + 5.0 ** 0.5) * k
        c = [(me.lower[d] + 0.5 * n) * h for d in range(3)]
        R = 0.25 * n * h
        X = np.stack([c[0] + R * np.cos(th) * np.sin(phi), c[1] + R * np.sin(th) * np.sin(phi), c[2] + R * np.cos(phi)], axis=1)
    if args.markers == "mixed":
        # BASELINE config C3: half of the markers uniform, half on a jittered spherical shell (radius n/4 cells, a few cells
        # thick) in the middle of the rank's patch
        half = N // 2
        k = np.arange(half, dtype=np.float64) + 0.5
        phi = np.arccos(1.0 - 2.0 * k / half)
        th = np.pi * (1.0 + 5.0 ** 0.5) * k
        c = [(me.lower[d] + 0.5 * n) * h for d in range(3)]
        R = (0.25 * n + 3.0 * (splitmix_unit(91, idx[:half]) - 0.5)) * h
        X[:half] = np.stack([c[0] + R * np.cos(th) * np.sin(phi), c[1] + R * np.sin(th) * np.sin(phi), c[2] + R * np.cos(phi)], axis=1)
    F = np.stack([2.0 * splitmix_unit(1 + d, idx) - 1.0 for d in range(3)], axis=1)
    # pinned host buffers for the e2e leg
    hX = torch.from_numpy(X).pin_memory()
    hF = torch.from_numpy(F).pin_memory()
    hU = torch.zeros((N, 3), dtype=torch.float64).pin_memory()
    hu, hf = [], []
    for a in range(3):
        shp = ib.side_shape(0, a)
        lin = [np.arange(shp[2 - d], dtype=np.float64) for d in range(3)]  # index along dim d
        coords = []
        for d in range(3):
            c = (me.lower[d] - g + lin[d] + (0.0 if d == a else 0.5)) * h
            coords.append(c)
        ua = (np.sin(2 * np.pi * coords[a]).reshape([-1 if d == a else 1 for d in (2, 1, 0)]) *
              np.cos(2 * np.pi * coords[(a + 1) % 3]).reshape([-1 if d == (a + 1) % 3 else 1 for d in (2, 1, 0)]))
        ua = np.broadcast_to(ua, shp).copy()
        t = torch.from_numpy(ua).pin_memory()
        hu.append(t)
        hf.append(torch.zeros(shp, dtype=torch.float64).pin_memory())
        ib.grid_upload("u", 0, a, t.numpy())
    ib.setPositions(hX.numpy())
    ib.setLData("F", hF.numpy())
    ctx.enable_timing(True)
    ib.beginDataRedistribution()
    ctx.synchronize()
    ib.beginDataRedistribution()  # the first call allocates (sort buffers, shadow columns): time a warm one
    ctx.synchronize()
    rebin_ms = ctx.last_ms(2)
    touched = ib.count_touched_dofs(KERNEL)

    # Multi-rank: the messages of the halo exchange are in flight while the tiles that do not touch the exchanged
    # regions are processed (ibk_*_part; include/ibk.h).
    part_ms = {"spread": 0.0, "interp": 0.0}  # multi-rank: device time of the tile kernels, both parts (last step)

    def timed_part(key, fn, part):
        if not part_ms.get("on"):
            return fn(part)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(part)
        e1.record(stream)
        part_ms.setdefault("events", []).append((key, e0, e1))

    def spread_part():
        if hx is None:
            ib.spreadForce(accumulate_halo=True)
        else:
            lib, hnd = ctx.lib, ctx.h
            ctx.check(lib.ibk_spread_begin(hnd))
            timed_part("spread", ib.spreadForcePart, 2)  # boundary tiles: everything the neighbours need
            hx.accumulate_post()                         # pack + start the messages
            timed_part("spread", ib.spreadForcePart, 1)  # interior tiles, overlapping the transfers
            ib.halo("f")
            hx.accumulate_finish()       # wait (on the stream) + add
            ctx.check(lib.ibk_spread_end(hnd))

    def interp_part():
        if hx is None:
            ib.interpolateVelocity(fill_halo=True)
        else:
            hx.fill_post()
            ib.halo("u")
            timed_part("interp", ib.interpolateVelocityPart, 1)  # interior tiles read no ghost cell
            hx.fill_finish()
            timed_part("interp", ib.interpolateVelocityPart, 2)

    def step_split():
        # every operation complete before the next starts; its exchange hidden behind its own interior tiles
        spread_part()
        interp_part()

    def step_pipelined():
        # Multi-rank: the two operations of a step work on different fields, so each exchange travels during the OTHER
        # operation's kernel and neither kernel is cut into a boundary and an interior part: the ghost values of u are on
        # their way while f is spread, the ghost contributions of f while U is interpolated.  Everything is complete when
        # the step ends (accumulate_finish + ibk_spread_end are its last calls).
        lib, hnd = ctx.lib, ctx.h
        hx.fill_post()                                   # u: pack + start the messages
        ctx.check(lib.ibk_spread_begin(hnd))
        timed_part("spread", ib.spreadForcePart, 0)      # all tiles, one launch
        hx.accumulate_post()                             # f: pack what the neighbours own + start the messages
        ib.halo("f")                                     # same-process ghost accumulation (after the pack)
        ib.halo("u")
        hx.fill_finish()                                 # u ghosts: wait (on the stream) + unpack
        timed_part("interp", ib.interpolateVelocityPart, 0)
        hx.accumulate_finish()                           # f: wait + add in ascending source-rank order
        ctx.check(lib.ibk_spread_end(hnd))

    pipelined = hx is not None and os.environ.get("IBK_BENCH_SEQUENCE", "pipelined") != "split"
    if pipelined:
        ctx.check(ctx.lib.ibk_comm_set_reserved_sms(ctx.h, int(os.environ.get("IBK_COMM_RESERVE_SMS", "0"))))

    def step():
        if hx is None:
            spread_part()
            interp_part()
        elif pipelined:
            step_pipelined()
        else:
            step_split()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    if os.environ.get("IBK_BENCH_PHASES") and hx is not None:
        # all ranks run the instrumented step (the exchange is collective); rank 0 prints
        lib, hnd = ctx.lib, ctx.h
        seq = [("spread_begin", lambda: ctx.check(lib.ibk_spread_begin(hnd))), ("spread boundary tiles", lambda: ib.spreadForcePart(2)),
               ("accumulate_post (pack+send)", hx.accumulate_post), ("spread interior tiles", lambda: ib.spreadForcePart(1)),
               ("halo_local f", lambda: ib.halo("f")), ("accumulate_finish (wait+unpack)", hx.accumulate_finish),
               ("spread_end", lambda: ctx.check(lib.ibk_spread_end(hnd))), ("fill_post (pack+send)", hx.fill_post),
               ("halo_local u", lambda: ib.halo("u")), ("interp interior tiles", lambda: ib.interpolateVelocityPart(1)),
               ("fill_finish (wait+unpack)", hx.fill_finish), ("interp boundary tiles", lambda: ib.interpolateVelocityPart(2))]
        if pipelined:
            seq = [("fill_post u (pack+send)", hx.fill_post), ("spread_begin", lambda: ctx.check(lib.ibk_spread_begin(hnd))),
                   ("spread, all tiles", lambda: ib.spreadForcePart(0)), ("accumulate_post f (pack+send)", hx.accumulate_post),
                   ("halo_local f", lambda: ib.halo("f")), ("halo_local u", lambda: ib.halo("u")),
                   ("fill_finish u (wait+unpack)", hx.fill_finish), ("interp, all tiles", lambda: ib.interpolateVelocityPart(0)),
                   ("accumulate_finish f (wait+add)", hx.accumulate_finish), ("spread_end", lambda: ctx.check(lib.ibk_spread_end(hnd)))]
        for rep in range(2):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)]
            evs[0].record(stream)
            for k, (_, fn) in enumerate(seq):
                fn()
                evs[k + 1].record(stream)
            torch.cuda.synchronize()
            if rank == 0 and rep == 1:
                for k, (name, _) in enumerate(seq):
                    print(f"[phases] {name:34s} {evs[k].elapsed_time(evs[k + 1]):7.3f} ms", file=sys.stderr)
                print(f"[phases] total {evs[0].elapsed_time(evs[-1]):7.3f} ms", file=sys.stderr)
        # the bare exchanges, nothing overlapped
        for name, post, fin in (("fill", hx.fill_post, hx.fill_finish), ("accumulate", hx.accumulate_post, hx.accumulate_finish)):
            for rep in range(3):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                barrier()
                e0.record(stream)
                post()
                e1.record(stream)
                fin()
                e2.record(stream)
                torch.cuda.synchronize()
            if rank == 0:
                print(f"[phases] bare {name}: post {e0.elapsed_time(e1):.3f} ms, finish {e1.elapsed_time(e2):.3f} ms", file=sys.stderr)
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    spread_ms, interp_ms = [], []
    launches0 = ctx.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    # per-kernel-group device times (CUDA events on the ctx stream inside libibk): a few extra steps,
    # read back one by one so the events of each step are still the "last" ones
    for _ in range(min(args.steps, 5)):
        part_ms["on"] = hx is not None
        part_ms["events"] = []
        step()
        ctx.synchronize()
        torch.cuda.synchronize()
        part_ms["on"] = False
        if hx is None:
            spread_ms.append(ctx.last_ms(0))
            interp_ms.append(ctx.last_ms(1))
        else:  # boundary + interior launches of each operation
            spread_ms.append(sum(a.elapsed_time(b) for k, a, b in part_ms["events"] if k == "spread"))
            interp_ms.append(sum(a.elapsed_time(b) for k, a, b in part_ms["events"] if k == "interp"))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    ms_per_step = total_ms / args.steps

    # ---- e2e: the same step through host buffers (pinned): X, F, u in; U, f out
    # The u upload and the f download run on the library's copy streams (ibk_grid_*_async): u is not needed
    # before the interpolation, so its upload overlaps the re-binning, the spread and the f download.
    C = __import__("ctypes")

    def e2e_step():
        ib.setLData("X", hX.numpy())  # new positions from the host ...
        ib.setLData("F", hF.numpy())
        for a in range(3):
            ib.grid_upload_async("u", 0, a, hu[a].numpy())
        ib.beginDataRedistribution()  # ... are re-binned on the device (wrap, cell, radix sort, permute)
        ib.grid_fill("f", 0.0)
        spread_part()
        for a in range(3):
            ib.grid_download_async("f", 0, a, hf[a].numpy())
        interp_part()
        ctx.check(ctx.lib.ibk_markers_download(ctx.h, 1, hU.numpy().ctypes.data_as(C.POINTER(C.c_double))))
        ib.transfers_wait()

    if args.e2e_breakdown and rank == 0:  # diagnostic: every phase alone, synchronised (stderr)
        def timed(label, fn):
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            ctx.synchronize()
            torch.cuda.synchronize()
            print(f"[e2e breakdown] {label:28s} {(time.perf_counter() - t) * 1e3:8.2f} ms", file=sys.stderr)
        for _ in range(2):
            timed("X upload", lambda: ib.setLData("X", hX.numpy()))
            timed("F upload", lambda: ib.setLData("F", hF.numpy()))
            timed("u upload (sync api)", lambda: [ib.grid_upload("u", 0, a, hu[a].numpy()) for a in range(3)])
            timed("u upload (async api)", lambda: [ib.grid_upload_async("u", 0, a, hu[a].numpy()) for a in range(3)])
            timed("rebin", ib.beginDataRedistribution)
            timed("f fill", lambda: ib.grid_fill("f", 0.0))
            timed("spread", spread_part)
            timed("f download (async api)", lambda: [ib.grid_download_async("f", 0, a, hf[a].numpy()) for a in range(3)])
            timed("interp", interp_part)
            timed("U download", lambda: ctx.check(ctx.lib.ibk_markers_download(
                ctx.h, 1, hU.numpy().ctypes.data_as(C.POINTER(C.c_double)))))
            timed("whole e2e step", e2e_step)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    grid_bytes = sum(int(np.prod(ib.side_shape(0, a))) for a in range(3)) * 8
    h2d = 2 * N * 3 * 8 + grid_bytes
    d2h = N * 3 * 8 + grid_bytes
    # ---- checks on the result of the last e2e step (hf = S[F] into a zeroed f, hU = J[u]); every rank takes part
    check = invariant_check(hf, hu, hF, hU, g, n, h, world, dist, torch)

    # ---- reduce over ranks (max time)
    if world > 1:
        t = torch.tensor([ms_per_step, e2e_s, float(np.mean(spread_ms)), float(np.mean(interp_ms))], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_s, sp_ms, in_ms = [float(v) for v in t.tolist()]
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    else:
        sp_ms, in_ms = float(np.mean(spread_ms)), float(np.mean(interp_ms))

    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks["hbm_gbs"])
        total_markers = N * world
        spread_bytes = 8 * (3 * N + 3 * N) + 16 * touched
        interp_bytes = 8 * (3 * N + 3 * N) + 8 * touched
        ach = spread_bytes / (sp_ms * 1e-3) / 1e9
        # DRAM traffic of the same kernels from the committed ncu pass (profiles/, not measured in this run)
        traffic, traffic_interp, traffic_src = None, None, None
        try:
            import glob
            newest = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_step_kernels_dram.json")))[-1]
            with open(newest) as f:
                prof = json.load(f)["per_kernel"]
            traffic = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in prof.items() if k.startswith("spread_"))
            traffic_interp = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in prof.items() if k.startswith("interp_"))
            traffic_src = f"profiles/{os.path.basename(newest)} (ncu dram__bytes_read.sum + dram__bytes_write.sum, same workload, 1 GPU)"
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_markers / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, args),
            "e2e": {"value": total_markers / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "what": "host X, F, u (pinned) -> device, re-bin, step, U and f -> host; all copies inside the timed region, u upload / f download on copy streams overlapping the kernels"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": f"spread_march_kernel<{KERNEL}> (+ dense bricks, fix-up)", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": traffic if args.config == "C5" else None, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes": spread_bytes, "launch_ms": sp_ms,
                         "interp": {"kernel": "interp_rot_kernel<IB_4,320>" if KERNEL == "IB_4" else f"interp_tile_kernel<3,{KERNEL}>", "achieved": interp_bytes / (in_ms * 1e-3) / 1e9,
                                    "frac": interp_bytes / (in_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": interp_bytes,
                                    "traffic": traffic_interp,
                                    "launch_ms": in_ms},
                         "touched_side_dofs": touched},
            "clocks": sampler.result(),
            "phases_ms": {"spread_kernels": sp_ms, "interp_kernels": in_ms, "halo_and_gaps": ms_per_step - sp_ms - in_ms, "step": ms_per_step,
                          "rebin": rebin_ms},
        }
        if world > 1:
            line["config"]["step"] += ("; multi-rank sequence: " + (
                "the u ghost exchange (NCCL send/recv over NVLink, libibk.so's communicator) travels during the spread kernel and the f "
                "ghost accumulation exchange during the interpolation kernel, both complete inside the step" if pipelined else
                "each operation's exchange overlaps its own interior tiles (boundary tiles first)"))
        line["check"] = check
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(n=args.cells, log2_markers=args.log2_markers)  # the N = 1 workload itself (what --impl reference runs)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not args.no_sample_parity:
            line["check"]["sample_vs_oracle"] = sample_parity(local_rank)
            line["check"]["ok"] = bool(line["check"]["ok"] and line["check"]["sample_vs_oracle"]["ok"])
        print(json.dumps(line), flush=True)
        if not line["check"]["ok"]:
            raise SystemExit("bench.py: result check FAILED: " + json.dumps(line["check"]))
    ib.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=512, help="cells per dimension per GPU")
    ap.add_argument("--log2-markers", type=int, default=23, help="log2 of the markers per GPU")
    ap.add_argument("--markers", default=None, choices=["uniform", "shell", "mixed"],
                    help="marker distribution: uniform (C5), a dense spherical shell (C2), half uniform + half jittered shell (C3)")
    ap.add_argument("--config", default="C5", choices=["C5", "C3", "C2"],
                    help="BASELINE.json configuration: C5 (default, the headline: IB_4, 512^3 + 2^23 uniform markers per GPU), "
                         "C3 (IB_6, 512^3, 2^23 markers uniform + clustered shell), C2 (IB_4, 256^3, 2^20 markers on a shell)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-breakdown", action="store_true", help="print the e2e phases, each synchronised, to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sample-parity", action="store_true", help="skip the 256^3 value-by-value comparison with the oracle (N = 1)")
    args = ap.parse_args()
    global KERNEL, METRIC
    if args.config == "C3":
        KERNEL, METRIC = "IB_6", "IB_6 spread+interp markers/sec"
        args.markers = args.markers or "mixed"
    elif args.config == "C2":
        args.markers = args.markers or "shell"
        if args.cells == 512 and args.log2_markers == 23:
            args.cells, args.log2_markers = 256, 20
    args.markers = args.markers or "uniform"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
