"""Index arithmetic of the 2 x 2 x 2 cluster write-out of spread_tile_kernel<..., CL = true> (ibk_spread.cu), replayed on the CPU.

Eight CTAs hold haloed blocks of (16 + 2M)^3 points for the eight 16^3 tiles of a 32^3 cluster tile.  Each CTA owns a share of
(16 + M)^3 points of the cluster's (32 + 2M)^3 points; the total at a share point is the sum over the blocks that hold it, in rank
order.  Checked here: the shares tile the cluster region exactly once, the kernel's sibling-coordinate formulas pick the right
points, and the chunked in-place dense repack of a share (reads of a chunk, barrier, writes of the chunk) never reads a value it
has already overwritten."""
import numpy as np
import pytest

TILE = 16


@pytest.mark.parametrize("M", [2, 4])
def test_shares_partition_the_cluster_region_and_sum_all_blocks(M):
    R, SH, E = TILE + 2 * M, TILE + M, 2 * TILE + 2 * M
    rng = np.random.default_rng(3)
    blocks = rng.standard_normal((8, R, R, R))          # [rank][z][y][x]
    # reference: drop every block into the cluster region at its offset 16 * (rank bit) and add up
    ref = np.zeros((E, E, E))
    for s in range(8):
        oz, oy, ox = TILE * ((s >> 2) & 1), TILE * ((s >> 1) & 1), TILE * (s & 1)
        ref[oz:oz + R, oy:oy + R, ox:ox + R] += blocks[s]
    got = np.full((E, E, E), np.nan)
    for crank in range(8):
        o = [M if (crank >> d) & 1 else 0 for d in range(3)]        # share origin inside the CTA's block (x, y, z)
        for q in range(SH ** 3):
            lx, ly, lz = q % SH + o[0], (q // SH) % SH + o[1], q // (SH * SH) + o[2]
            both = ((lx < 2 * M) if crank & 1 else (lx >= TILE)) | (((ly < 2 * M) if crank & 2 else (ly >= TILE)) << 1) | \
                   (((lz < 2 * M) if crank & 4 else (lz >= TILE)) << 2)
            v = 0.0
            for sr in range(8):
                diff = sr ^ crank
                if diff & ~both:
                    continue
                cx = lx + ((TILE if crank & 1 else -TILE) if diff & 1 else 0)
                cy = ly + ((TILE if crank & 2 else -TILE) if diff & 2 else 0)
                cz = lz + ((TILE if crank & 4 else -TILE) if diff & 4 else 0)
                assert 0 <= cx < R and 0 <= cy < R and 0 <= cz < R
                v += blocks[sr][cz, cy, cx]
            gx, gy, gz = lx + TILE * (crank & 1), ly + TILE * ((crank >> 1) & 1), lz + TILE * ((crank >> 2) & 1)
            assert np.isnan(got[gz, gy, gx]), "a point belongs to exactly one share"
            got[gz, gy, gx] = v
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)


@pytest.mark.parametrize("crank", range(8))
def test_chunked_in_place_repack_is_safe(crank):
    M, NT = 2, 256
    R, RX, SH = TILE + 2 * M, TILE + 2 * M, TILE + M
    o0, o1, o2 = (M if crank & 1 else 0), (M if crank & 2 else 0), (M if crank & 4 else 0)
    acc = np.arange(R * R * RX, dtype=np.float64)
    want = acc.reshape(R, R, RX)[o2:o2 + SH, o1:o1 + SH, o0:o0 + SH].reshape(-1).copy()
    for q0 in range(0, SH ** 3, NT):
        q = np.arange(q0, min(q0 + NT, SH ** 3))
        src = ((q // (SH * SH) + o2) * R + (q // SH) % SH + o1) * RX + q % SH + o0
        assert np.all(src >= q)
        val = acc[src].copy()      # all reads of the chunk ...
        acc[q] = val               # ... then (after the barrier) all its writes
    np.testing.assert_array_equal(acc[:SH ** 3], want)
