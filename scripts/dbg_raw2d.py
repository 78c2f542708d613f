import sys, numpy as np
sys.path.insert(0, '.')
from ibamr_b200 import api
from oracle import oracle as orc
ndim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kernel = sys.argv[2] if len(sys.argv) > 2 else "IB_4"
g = orc.min_ghost_width(kernel)
ilower = (3, -2, 5)[:ndim]; iupper = (23, 15, 18)[:ndim]; nugc = (g, g + 1, g)[:ndim]
dx = (0.05, 0.04, 0.0625)[:ndim]; x_lower = (-0.3, 0.1, 0.25)[:ndim]
x_upper = tuple(x_lower[d] + dx[d] * (iupper[d] - ilower[d] + 1) for d in range(ndim))
depth = 2
shape = (depth,) + tuple(reversed([iupper[d] - ilower[d] + 1 + 2 * nugc[d] for d in range(ndim)]))
rng = np.random.default_rng(1)
u = rng.standard_normal(shape); n = 400
X = np.stack([rng.uniform(x_lower[d] - 4 * dx[d], x_upper[d] + 4 * dx[d], n) for d in range(ndim)], axis=1)
idx = np.arange(n, dtype=np.int32)
V = np.zeros((n, depth)); Vr = np.zeros((n, depth))
api.raw_interp_host(kernel, ndim, dx, x_lower, x_upper, depth, ilower, iupper, nugc, u, idx, None, X, V)
orc.interp_raw(kernel, ndim, dx, x_lower, depth, ilower, iupper, nugc, u, idx, np.zeros(n * ndim), X, Vr)
print("interp err", np.abs(V - Vr).max())
