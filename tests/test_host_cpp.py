"""The C++ host mirror (ibamr_b200/host/*.h + the Fortran link-substitution shim) compiles against the
C ABI, answers the LEInteractor static queries, refuses to compute without a GPU (CPU test), and reproduces
the oracle through LEInteractor::interpolate/spread and IBMethodB200 (GPU test)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host_cpp", "_build")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="module")
def driver():
    from ibamr_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "driver")
    src = os.path.join(ROOT, "tests", "host_cpp", "driver.cpp")
    libdir = os.path.join(ROOT, "ibamr_b200")
    deps = [src, _lib.LIB_PATH] + [os.path.join(libdir, "host", f) for f in os.listdir(os.path.join(libdir, "host"))]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run([CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, src, "-L" + libdir, "-libk",
                        "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
        # the Fortran-symbol shim must compile and define every entry point (21 kernels x interp/spread x 2d/3d)
        obj = os.path.join(BUILD, "shim.o")
        subprocess.run([CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-c", "-o", obj,
                        os.path.join(libdir, "host", "ibk_fortran_shim.cpp")], check=True, capture_output=True, text=True)
    return exe


def test_cpp_static_queries_and_refusal(driver):
    out = subprocess.run([driver, "--static"], capture_output=True, text=True, check=True).stdout
    assert "IB_4 known=1 stencil=4 ghosts=3" in out
    assert "IB_6 known=1 stencil=6 ghosts=4" in out
    assert "BSPLINE_3 known=1 stencil=4 ghosts=3" in out
    assert "PIECEWISE_LINEAR known=1 stencil=2 ghosts=2" in out
    assert "IB_7 known=0" in out and "unknown kernel: error raised" in out
    import torch
    if not torch.cuda.is_available():
        assert "ctx_create rc=-2" in out  # IBK_ERR_CUDA: no CPU fallback


def test_fortran_shim_exports_all_symbols(driver):
    out = subprocess.run(["nm", os.path.join(BUILD, "shim.o")], capture_output=True, text=True, check=True).stdout
    for k in ("piecewise_linear", "ib_4", "ib_6", "bspline_3", "bspline_4", "ib_3", "bspline_5", "bspline_6", "piecewise_cubic", "ib_5", "piecewise_constant",
              "composite_bspline_32", "composite_bspline_23", "composite_bspline_43", "composite_bspline_34", "composite_bspline_54",
              "composite_bspline_45", "composite_bspline_65", "composite_bspline_56", "discontinuous_linear", "ib_4_w8"):
        for op in ("interp", "spread"):
            for d in ("2d", "3d"):
                assert f" T lagrangian_{k}_{op}{d}_" in out


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["IB_4", "BSPLINE_3"])
def test_cpp_host_mirror_vs_oracle(driver, kernel, tmp_path):
    from oracle import oracle as orc
    from tests.util import splitmix64_unit
    n, N = 24, 3000
    g = orc.min_ghost_width(kernel)
    level = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0,) * 3, (n - 1,) * 3)], (g,) * 3)
    pg = level.patch_geom(0)
    X = np.stack([splitmix64_unit(5 + d, np.arange(N)) for d in range(3)], axis=1)
    F = np.stack([2 * splitmix64_unit(9 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    u = []
    ncell = (n,) * 3
    for a in range(3):
        c = pg.side_coords(a)
        u.append(np.ascontiguousarray(np.sin(2 * np.pi * c[a]) * np.cos(2 * np.pi * c[(a + 1) % 3])))
    case, outp = tmp_path / "case.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("iii", n, g, N))
        f.write(kernel.encode().ljust(32, b"\0"))
        f.write(X.tobytes())
        f.write(F.tobytes())
        for a in range(3):
            f.write(u[a].tobytes())
    r = subprocess.run([driver, str(case), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(outp, dtype=np.float64)
    sizes = [u[a].size for a in range(3)]
    off = 0
    Q = raw[off:off + 3 * N].reshape(N, 3); off += 3 * N
    fB3 = []
    for a in range(3):
        fB3.append(raw[off:off + sizes[a]].reshape(u[a].shape)); off += sizes[a]
    U = raw[off:off + 3 * N].reshape(N, 3); off += 3 * N
    fB1 = []
    for a in range(3):
        fB1.append(raw[off:off + sizes[a]].reshape(u[a].shape)); off += sizes[a]

    def rel(a, b):
        return np.max(np.abs(a - b)) / np.max(np.abs(b))

    # seam B3: position-only LEInteractor calls (no periodic images)
    Qo = orc.side_interp_positions(kernel, pg, u, X)
    assert rel(Q, Qo) <= 1e-12
    fo = [np.zeros_like(a) for a in u]
    orc.side_spread_positions(kernel, pg, fo, X, F)
    for a in range(3):
        assert rel(fB3[a], fo[a]) <= 1e-12
    # seam B1: the periodic level against the reference model (ghost-box lists with periodic shifts)
    ref = orc.bin_level(level, X)
    lst = ref["patches"][0]
    fr = [np.zeros_like(a) for a in u]
    orc.side_spread(kernel, pg, fr, X, F, lst["all_idx"], lst["all_shift"])
    for a in range(3):
        sl = tuple(slice(g, s - g) for s in fr[a].shape)
        assert rel(fB1[a][sl], fr[a][sl]) <= 1e-12
    # u given to the level has analytic (periodic) ghosts already, so U must equal the position-only result
    assert rel(U, Qo) <= 1e-12
    # N1 through the C++ mirror: computeLagrangianForce (ring of springs + target points) and forwardEulerStep
    Fl = raw[off:off + 3 * N].reshape(N, 3); off += 3 * N
    Xnew = raw[off:off + 3 * N].reshape(N, 3); off += 3 * N
    assert raw[off] == 1.0, "LDataB200: lazy host mirror (read, modify + restore, refetch after a kernel)"
    off += 1
    assert off == raw.size
    Xw, _ = orc.wrap_positions(X, level.x_lower, level.x_upper, level.periodic)
    Xw = Xw.reshape(N, 3)
    ti = np.arange(0, N, 7)
    Fo = orc.lagrangian_force(3, Xw, U, springs=(np.arange(N), (np.arange(N) + 1) % N, np.full(N, 1.5), np.full(N, 0.01)),
                              targets=(ti, np.full(len(ti), 2.0), np.full(len(ti), 0.25), np.full((len(ti), 3), 0.5)))
    assert rel(Fl, Fo) <= 1e-14
    assert np.array_equal(Xnew, Xw + 0.01 * U)


def test_cpp_structure_reader(driver):
    """IBStandardInitializerB200 (host C++, no GPU) on the reference's sample structure files."""
    gold = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([driver, "--structure", os.path.join(gold, "curve2d_64"), "2"], capture_output=True, text=True, check=True).stdout
    assert "vertices=304 springs=304 beams=0 targets=0 anchors=0" in out
    assert "last spring=0,303,193.53241079974475,0" in out
    out = subprocess.run([driver, "--structure", os.path.join(gold, "fila_256"), "2"], capture_output=True, text=True, check=True).stdout
    assert "vertices=201 springs=200 beams=199 targets=1 anchors=0" in out and "X0=4.5,14.75" in out
    r = subprocess.run([driver, "--structure", os.path.join(gold, "no_such_structure"), "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot find required vertex file" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["IB_4", "IB_6"])
def test_cpp_two_ranks_with_reference_signatures(driver, kernel, tmp_path):
    """Seams B1 / B2 across ranks from C++: two IBMethodB200 objects of one process are the two ranks of
    a loopback communicator of libibk.so; spreadForce / interpolateVelocity are called with the reference's signatures
    (data indices bound to host SideData, schedules as stand-ins), the halo plan, messages and unpack-adds all happen behind
    the C ABI.  A third, single-rank object goes through LDataManagerB200::spread (with node weights, F * ds on the device)
    and ::interp.  Everything against the oracle's model of the reference."""
    from oracle import oracle as orc
    from tests.util import splitmix64_unit
    n, N = 32, 4000
    g = orc.min_ghost_width(kernel)
    boxes = [((0, 0, 0), (n // 2 - 1, n - 1, n - 1)), ((n // 2, 0, 0), (n - 1, n - 1, n - 1))]
    level2 = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), boxes, (g,) * 3)
    level1 = orc.Level(3, (0,) * 3, (n,) * 3, (0.0,) * 3, (1.0,) * 3, (1, 1, 1), [((0,) * 3, (n - 1,) * 3)], (g,) * 3)
    pg1 = level1.patch_geom(0)
    X = np.stack([splitmix64_unit(5 + d, np.arange(N)) for d in range(3)], axis=1)
    F = np.stack([2 * splitmix64_unit(9 + d, np.arange(N)) - 1 for d in range(3)], axis=1)
    u = []
    for a in range(3):
        c = pg1.side_coords(a)
        u.append(np.ascontiguousarray(np.sin(2 * np.pi * c[a]) * np.cos(2 * np.pi * c[(a + 1) % 3]) + 0 * c[0] + 0 * c[1] + 0 * c[2]))
    case, outp = tmp_path / "case2.bin", tmp_path / "out2.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("iii", n, g, N))
        f.write(kernel.encode().ljust(32, b"\0"))
        f.write(X.tobytes())
        f.write(F.tobytes())
        for a in range(3):
            f.write(u[a].tobytes())
    r = subprocess.run([driver, "--two-ranks", str(case), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(outp, dtype=np.uint8)
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a

    def rel(a, b):
        return np.max(np.abs(a - b)) / np.max(np.abs(b))

    ref = orc.bin_level(level2, X)
    for rk in range(2):
        cnt = int(take(np.int64, 1)[0])
        ids = take(np.int64, cnt)
        assert np.array_equal(ids, np.nonzero(ref["owner"] == rk)[0])
        U = take(np.float64, 3 * cnt).reshape(cnt, 3)
        pg = level2.patch_geom(rk)
        f_got = [take(np.float64, int(np.prod(pg.side_shape(a)))).reshape(pg.side_shape(a)) for a in range(3)]
        lst = ref["patches"][rk]
        ii = lst["all_idx"][lst["interior_mask"]]
        sh = lst["all_shift"].reshape(-1, 3)[lst["interior_mask"]]
        ur = []
        for a in range(3):  # the rank's u with analytic (periodic) ghosts: what the exchange must have produced
            c = pg.side_coords(a)
            ur.append(np.ascontiguousarray(np.sin(2 * np.pi * c[a]) * np.cos(2 * np.pi * c[(a + 1) % 3]) + 0 * c[0] + 0 * c[1] + 0 * c[2]))
        U_ref = orc.side_interp(kernel, pg, ur, X, ii, sh.reshape(-1))
        assert rel(U, U_ref[ids]) <= 1e-12
        f_ref = [np.zeros(pg.side_shape(a)) for a in range(3)]
        orc.side_spread(kernel, pg, f_ref, X, F, lst["all_idx"], lst["all_shift"])
        for a in range(3):
            sl = tuple(slice(g, s - g) for s in f_ref[a].shape)
            assert np.max(np.abs(f_got[a][sl] - 0.25 - f_ref[a][sl])) <= 1e-12 * np.max(np.abs(f_ref[a][sl]))
    # seam B2: spread of F * ds, interp into an auxiliary LData
    ds = 0.5 + 0.001 * (np.arange(N) % 100)
    ref1 = orc.bin_level(level1, X)
    lst = ref1["patches"][0]
    fr = [np.zeros_like(a) for a in u]
    orc.side_spread(kernel, pg1, fr, X, F * ds[:, None], lst["all_idx"], lst["all_shift"])
    for a in range(3):
        got = take(np.float64, u[a].size).reshape(u[a].shape)
        sl = tuple(slice(g, s - g) for s in fr[a].shape)
        assert rel(got[sl], fr[a][sl]) <= 1e-12
    Ua = take(np.float64, 3 * N).reshape(N, 3)
    assert rel(Ua, orc.side_interp_positions(kernel, pg1, u, X)) <= 1e-12
    assert off == raw.size


@pytest.mark.gpu
def test_cpp_ldata_restart_round_trip(driver, tmp_path):
    """N2, restart part (missing 4): LData::putToDatabase / LData(Pointer<Database>) (LData.cpp:99-130, 186-209)
    over device-resident columns: X, U, F written to files, a second level rebuilt from the files alone; the columns come
    back bit for bit and the spread after the restart equals the spread before it bit for bit."""
    r = subprocess.run([driver, "--restart", str(tmp_path / "ldata")], capture_output=True, text=True)
    assert r.returncode == 0 and "restart ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_two_level_prolong_and_coarsen(driver):
    """N3 through the C++ seam: LDataManagerB200::setCoarserLevel + the schedule arguments of spread / interp
    (LDataManager.cpp:611-614, 728-734) over two levels resident on one device."""
    r = subprocess.run([driver, "--amr"], capture_output=True, text=True)
    assert r.returncode == 0 and "amr ok" in r.stdout, r.stdout + r.stderr
