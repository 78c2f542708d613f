"""The CPU restatement of IBStandardForceGen's spring / beam / target-point loops (oracle/le_force.c) against
hand-computed cases and the invariants of the force laws (the reference's tests hold no force golden output;
SURVEY.md 8(f) N1)."""
import os

import numpy as np

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_two_node_spring_by_hand():
    X = np.array([[0.0, 0.0], [3.0, 4.0]])
    F = orc.lagrangian_force(2, X, np.zeros_like(X), springs=([0], [1], [2.0], [1.0]))
    # R = 5, T = 2 (5 - 1) = 8, direction (0.6, 0.8) on the master, opposite on the slave
    assert np.allclose(F, [[4.8, 6.4], [-4.8, -6.4]], rtol=1e-15)
    # coincident nodes: skipped (IBStandardForceGen.cpp:868)
    assert np.all(orc.lagrangian_force(2, np.zeros((2, 2)), np.zeros((2, 2)), springs=([0], [1], [2.0], [1.0])) == 0.0)


def test_beam_and_target_by_hand():
    X = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 0.0], [2.0, 0.0, 0.0]])
    F = orc.lagrangian_force(3, X, np.zeros_like(X), beams=([1], [2], [0], [0.5], [[0.0, 0.0, 0.0]]))
    # K (X_next + X_prev - 2 X_mastr) = 0.5 * (0, -2, 0) = (0, -1, 0): mastr += 2F, the others -= F
    assert np.allclose(F, [[0, 1, 0], [0, -2, 0], [0, 1, 0]])
    U = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 2.0, 0.0]])
    F = orc.lagrangian_force(3, X, U, targets=([0, 2], [10.0, 4.0], [0.5, 0.25], [[0.1, 0, 0], [2, 0, 1.0]]))
    assert np.allclose(F, [[10 * 0.1 - 0.5, 0, 0], [0, 0, 0], [0, -0.5, 4.0]])


def test_closed_ring_of_the_reference_example_is_in_equilibrium_of_total_force():
    """curve2d_64: internal forces (springs) sum to zero; zero-rest-length springs on a closed curve give
    F_l = K (X_{l+1} - 2 X_l + X_{l-1})."""
    X = orc.read_vertex_file(os.path.join(GOLD, "curve2d_64.vertex"), 2)
    m, s, k, r, _ = orc.read_spring_file(os.path.join(GOLD, "curve2d_64.spring"), len(X))
    F = orc.lagrangian_force(2, X, np.zeros_like(X), springs=(m, s, k, r))
    assert np.max(np.abs(F.sum(axis=0))) < 1e-10 * np.max(np.abs(F))
    lap = k[0] * (np.roll(X, -1, axis=0) - 2 * X + np.roll(X, 1, axis=0))
    assert np.max(np.abs(F - lap)) < 1e-12 * np.max(np.abs(lap))
