"""Measures what the host<->device link of the GPU box gives: flat and pitched (cudaMemcpy2D) copies of one
side-centred component of the C5 shard, each direction alone and both at once.  Used to size the e2e leg."""
import time
import torch

n = (521, 520, 520)  # one component incl. ghosts (x fastest)
rows, rowlen = n[1] * n[2], n[0]
pitch = 528
h = torch.empty(rows * rowlen, dtype=torch.float64).pin_memory()
h2 = torch.empty(rows * rowlen, dtype=torch.float64).pin_memory()
d = torch.empty(rows * pitch, dtype=torch.float64, device="cuda")
d2 = torch.empty(rows * pitch, dtype=torch.float64, device="cuda")
gb = rows * rowlen * 8 / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def flat_h2d():
    with torch.cuda.stream(s1):
        d[: rows * rowlen].copy_(h, non_blocking=True)


def flat_d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2[: rows * rowlen], non_blocking=True)


def pitched_h2d():
    with torch.cuda.stream(s1):
        d.view(rows, pitch)[:, :rowlen].copy_(h.view(rows, rowlen), non_blocking=True)


def pitched_d2h():
    with torch.cuda.stream(s2):
        h2.view(rows, rowlen).copy_(d2.view(rows, pitch)[:, :rowlen], non_blocking=True)


def both_flat():
    flat_h2d()
    flat_d2h()


for name, fn, vol in (("flat h2d", flat_h2d, gb), ("flat d2h", flat_d2h, gb), ("pitched h2d", pitched_h2d, gb),
                      ("pitched d2h", pitched_d2h, gb), ("flat both directions", both_flat, 2 * gb)):
    t = timed(fn)
    print(f"{name:24s} {vol / t:7.1f} GB/s  ({t * 1e3:.1f} ms for {vol:.2f} GB)")
