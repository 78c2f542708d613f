"""The geometry spread_march_kernel (ibamr_b200/csrc/ibk_spread.cu) relies on for a race-free, bit-reproducible spread, replayed
on the CPU for every reach M the kernel serves:

 * bricks (4 cells) that are worked on at the same time have disjoint footprints (brick -+ M cells): inside a row the bricks NC
   apart (chains of a sub-phase), between rows the rows NC apart (row colours), NC = ceil((4 + 2 M) / 4);
 * a row of colour ph + 1 overlaps only the rows next to it, so a consumer pair may start its odd row once its own even row
   and the NEXT pair's even row are done (the neighbour-only barrier), and of the other half's bricks only the upper half's
   first one is a neighbour of a brick of the next sub-phase (the arrive / sync split of the sub-phase barrier);
 * the ring of z planes: while layer s accumulates into its FZ = 4 + 2 M planes, the four planes layer s - 1 hands to the
   flusher are none of them, and no two live planes share a ring slot (NRING = FZ + 4);
 * march tiles 2 apart per dimension (the eight colours) have disjoint accumulators, and a tile's up to 26 neighbours of a lower
   colour hold smaller tickets (colour-major order): waiting for them cannot deadlock and fixes the order of the additions."""
import itertools

import pytest

BRICK, MARCH_CELLS, MARCH_BR = 4, 32, 8


def cfg(M):
    FZ = BRICK + 2 * M
    return dict(M=M, FZ=FZ, NRING=FZ + BRICK, NC=(BRICK + 2 * M + BRICK - 1) // BRICK)


def footprint(b, M):
    """cells a brick's markers can touch along one dimension"""
    return set(range(BRICK * b - M, BRICK * b + BRICK + M))


@pytest.mark.parametrize("M", [1, 2, 3])
def test_concurrent_bricks_and_rows_are_disjoint(M):
    NC = cfg(M)["NC"]
    for sp in range(NC):  # the chains of a sub-phase: bricks sp, sp + NC, ... of a row
        chain = list(range(sp, MARCH_BR, NC))
        for a, b in itertools.combinations(chain, 2):
            assert not (footprint(a, M) & footprint(b, M))
    for ph in range(NC):  # the rows of a colour
        rows = list(range(ph, MARCH_BR, NC))
        for a, b in itertools.combinations(rows, 2):
            assert not (footprint(a, M) & footprint(b, M))
    # bricks / rows that are NOT NC apart do overlap (the colouring is not wasteful)
    assert footprint(0, M) & footprint(NC - 1, M)


def test_fast4_barriers_cover_every_dependency():
    """M = 2 (the 4-point kernels): NC = 2, a pair of warps per row, two chains per warp."""
    M, NC = 2, 2
    # rows: pair s owns rows 2 s (colour 0) and 2 s + 1 (colour 1); row 2 s + 1 overlaps exactly rows 2 s and 2 s + 2
    for s in range(MARCH_BR // 2):
        odd = 2 * s + 1
        overlapping = {r for r in range(MARCH_BR) if r != odd and footprint(r, M) & footprint(odd, M)}
        assert overlapping == {r for r in (2 * s, 2 * s + 2) if r < MARCH_BR}
        # ... which belong to pair s (its own: the two warps of the pair synchronise) and pair s + 1 (it announces, pair s waits)
        assert {r // 2 for r in overlapping} <= {s, s + 1}
    # sub-phases inside a row: half h of the pair walks bricks sp + 2 (2 h + c), c = 0, 1
    bricks = lambda sp, half: [sp + NC * (2 * half + c) for c in range(2)]
    assert bricks(0, 0) == [0, 2] and bricks(0, 1) == [4, 6] and bricks(1, 0) == [1, 3] and bricks(1, 1) == [5, 7]
    # which bricks of sub-phase 0 does a brick of sub-phase 1 overlap, and whose are they?
    for half in (0, 1):
        needs_other_half = set()
        for b in bricks(1, half):
            for o in bricks(0, 1 - half):
                if footprint(b, M) & footprint(o, M):
                    needs_other_half.add((b, o))
        if half == 0:
            assert needs_other_half == {(3, 4)}  # the lower half waits for the upper half's FIRST brick only
        else:
            assert needs_other_half == set()     # the upper half needs nothing from the lower one: it announces and goes on


@pytest.mark.parametrize("M", [1, 2, 3])
def test_plane_ring(M):
    c = cfg(M)
    FZ, NRING = c["FZ"], c["NRING"]
    for s in range(1, 40):
        live = set(range(BRICK * s, BRICK * s + FZ))            # planes layer s accumulates into (relative to the chunk's first plane)
        flushing = set(range(BRICK * (s - 1), BRICK * s))       # the four planes layer s - 1 has made final
        assert not (live & flushing)
        slots = [q % NRING for q in sorted(live | flushing)]
        assert len(set(slots)) == len(slots) == NRING           # every ring slot holds exactly one plane
        # the planes layer s + 1 will need beyond those of layer s are exactly the slots the flusher frees (after zeroing them)
        new = set(range(BRICK * s + FZ, BRICK * (s + 1) + FZ))
        assert {q % NRING for q in new} == {q % NRING for q in flushing}


@pytest.mark.parametrize("M", [1, 2, 3])
def test_tile_colours_and_ticket_order(M):
    XO = M & 1
    R = MARCH_CELLS + 2 * M

    def block(col):  # accumulator extent of march tile `col` along x (the widest: + the alignment column)
        return set(range(MARCH_CELLS * col - M - XO, MARCH_CELLS * col - M - XO + ((R + XO + 1) & ~1)))
    for a in range(0, 6):
        assert not (block(a) & block(a + 2))   # same colour: disjoint
        assert block(a) & block(a + 1)         # neighbours overlap: the order between them has to be fixed
    # colour-major tickets: colour = x parity + 2 y parity + 4 z parity; every neighbour of a LOWER colour has a smaller ticket
    nm = (5, 4, 3)
    tiles = list(itertools.product(range(nm[0]), range(nm[1]), range(nm[2])))
    colour = lambda t: (t[0] & 1) + 2 * (t[1] & 1) + 4 * (t[2] & 1)
    order = sorted(tiles, key=lambda t: (colour(t), t[2] // 2, t[1] // 2, t[0] // 2))
    ticket = {t: k for k, t in enumerate(order)}
    for t in tiles:
        for d in itertools.product((-1, 0, 1), repeat=3):
            n = tuple(t[i] + d[i] for i in range(3))
            if n == t or n not in ticket:
                continue
            assert colour(n) != colour(t)      # no two neighbours share a colour: one of them always waits for the other
            if colour(n) < colour(t):
                assert ticket[n] < ticket[t]
