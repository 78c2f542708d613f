"""The bank arithmetic behind interp_rot_kernel (ibamr_b200/csrc/ibk_interp.cu), checked on the CPU.

The kernel's claim: in the staged 22 x 20 x 20 fp64 box, a lane that visits the 16 (j, k) rows of its marker's 4^3 stencil in the
rotated order q + r (q = 4 k + j, r = 3 ((target - b) mod 16) / 2 mod 8) lands on 8-byte bank `target + 6 t + i` at static step
(t, i), whatever its box origin b is -- provided b and target have the same parity.  With 8 even-origin and 8 odd-origin markers
per half-warp and target = 2 (l & 7) + (l >> 3), the 16 lanes of a half-warp therefore hit 16 different banks at every one of the
64 loads.  This test replays exactly that arithmetic (including the carry from j into k) for random origins."""
import numpy as np

S, SX, W = 20, 22, 4


def _banks_of_half_warp(x0, y0, z0):
    b = (z0 * S + y0) * SX + x0                       # element index of the stencil origin in the box
    lane = np.arange(16)
    grp = lane >> 3
    assert np.all((b & 1) == grp), "lanes 0-7 hold even origins, lanes 8-15 odd ones"
    tgt = 2 * (lane & 7) + grp
    r = (3 * (((tgt - b) & 15) >> 1)) & 7
    rj, rk = r & 3, r >> 2
    out = []
    for k in range(W):
        for j in range(W):
            jj = (j + rj) & 3
            carry = (j + rj) >= 4
            kk = np.where(carry, (k + 1 + rk) & 3, (k + rk) & 3)
            row = b + jj * SX + kk * (S * SX)
            for i in range(W):
                out.append(((row + i) % 16, jj, kk))
    return out


def test_rotated_rows_are_a_permutation_and_conflict_free():
    rng = np.random.default_rng(7)
    for _ in range(300):
        # 8 markers with an even x origin, 8 with an odd one; any y, z origin inside the box
        x0 = np.concatenate([2 * rng.integers(0, 9, 8), 2 * rng.integers(0, 9, 8) + 1])
        y0 = rng.integers(0, S - W + 1, 16)
        z0 = rng.integers(0, S - W + 1, 16)
        steps = _banks_of_half_warp(x0, y0, z0)
        assert len(steps) == 64
        for banks, _, _ in steps:
            assert len(set(banks.tolist())) == 16          # one wavefront per half-warp load
        # every lane still visits each of its 16 rows exactly once
        for lane in range(16):
            rows = {(int(jj[lane]), int(kk[lane])) for _, jj, kk in steps[::W]}
            assert len(rows) == 16


def test_unrotated_gather_conflicts():
    """The same origins without the rotation: about three wavefronts per half-warp load (what ncu showed: 2.9)."""
    rng = np.random.default_rng(8)
    worst = []
    for _ in range(300):
        b = (rng.integers(0, 17, 16) * S + rng.integers(0, 17, 16)) * SX + rng.integers(0, 18, 16)
        worst.append(np.bincount(b % 16, minlength=16).max())
    assert 2.5 < np.mean(worst) < 3.6
