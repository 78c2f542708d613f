"""N2: the C-ABI readers of IBStandardInitializer's structure files (host code in libibk.so, no
GPU needed) against (a) the reference's own sample files, whose content is known in closed form from the scripts
that generated them, and (b) the numpy restatement in oracle/oracle.py; plus the reference's error behaviour
(src/IB/IBStandardInitializer.cpp:184-294, 297-528, 766-1002, 1322-1517, 1520-1643)."""
import os

import numpy as np
import pytest

from ibamr_b200 import api
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_curve2d_64_matches_its_generator():
    """examples/IB/explicit/ex1/generate_curve2d.m:19-20,44-45,57-72: ellipse vertices, one spring per vertex to
    its successor with K = kappa / ds, rest length 0; the last edge (303, 0) is stored smaller index first."""
    init = api.IBStandardInitializer(2, [os.path.join(GOLD, "curve2d_64")])
    X = init.positions()
    assert X.shape == (304, 2)
    alpha, beta = 0.25 ** 2 / 0.35, 0.35
    th = 2.0 * np.pi * np.arange(304) / 304
    assert np.max(np.abs(X - np.stack([0.5 + alpha * np.cos(th), 0.5 + beta * np.sin(th)], axis=1))) < 5e-16
    m, s, k, r, f = init.springs[0]
    assert len(m) == 304 and np.all(f == 0) and np.all(r == 0.0)
    assert np.array_equal(m[:303], np.arange(303)) and np.array_equal(s[:303], np.arange(1, 304))
    assert (m[303], s[303]) == (0, 303)
    assert np.all(k == 1.9353241079974475e+02)
    assert len(init.beams[0][0]) == 0 and len(init.targets[0][0]) == 0 and len(init.anchors[0]) == 0


def test_fila_256_with_comments_and_all_file_kinds():
    """examples/IB/explicit/ex3/fila_256.*: every line carries a '#' comment; 201 vertices, 200 springs,
    199 beams (prev curr next), 1 target point without a damping field."""
    init = api.IBStandardInitializer(2, [os.path.join(GOLD, "fila_256")])
    assert init.positions().shape == (201, 2)
    assert np.allclose(init.positions()[0], [4.5, 14.75])
    m, s, k, r, f = init.springs[0]
    assert len(m) == 200 and np.array_equal(m, np.arange(200)) and np.array_equal(s, np.arange(1, 201))
    assert np.all(r == 1.4999999999999999e-02)
    pv, cu, nx, bend, curv = init.beams[0]
    assert len(cu) == 199 and np.array_equal(pv, np.arange(199)) and np.array_equal(cu, np.arange(1, 200))
    assert np.array_equal(nx, np.arange(2, 201)) and np.all(curv == 0.0)
    ti, tk, te = init.targets[0]
    assert ti.tolist() == [0] and tk.tolist() == [5.0e6] and te.tolist() == [0.0]
    # the numpy restatement reads the same
    om, os_, ok, orr, of = orc.read_spring_file(os.path.join(GOLD, "fila_256.spring"), 201)
    assert np.array_equal(om, m) and np.array_equal(os_, s) and np.array_equal(ok, k) and np.array_equal(orr, r)
    op, oc, on, ob, ocv = orc.read_beam_file(os.path.join(GOLD, "fila_256.beam"), 201, 2)
    assert np.array_equal(op, pv) and np.array_equal(oc, cu) and np.array_equal(on, nx) and np.array_equal(ob, bend)
    assert np.array_equal(orc.read_vertex_file(os.path.join(GOLD, "fila_256.vertex"), 2), init.positions())


def test_two_structures_are_offset(tmp_path):
    """Vertex numbers of the second structure are shifted by the size of the first (:203-210, :474-477)."""
    init = api.IBStandardInitializer(2, [os.path.join(GOLD, "curve2d_64"), os.path.join(GOLD, "fila_256")])
    assert init.n_vertices == 304 + 201
    assert init.springs[1][0][0] == 304 and init.springs[1][1][0] == 305
    assert init.targets[1][0].tolist() == [304]
    assert init.beams[1][1][0] == 305


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)[:-len(name.split(".")[-1]) - 1]


def test_duplicates_defaults_and_optional_fields(tmp_path):
    base = _write(tmp_path, "s.vertex", "4 ! four vertices\n0 0\n1 0 % trailing\n1 1\n0 1\n")
    _write(tmp_path, "s.spring", "4\n1 0 2.0 0.5\n0 1 9.0 9.0   # duplicate of the first, skipped\n2 3 1.0 0.0 0\n3 0 4.0 0.25 0 7.5\n")
    _write(tmp_path, "s.beam", "2\n0 1 2 3.0 0.1 0.2\n0 1 2 5.0 ! duplicate (curr, next, prev)\n")
    _write(tmp_path, "s.target", "3\n2 10.0 0.5\n2 99.0\n1 4.0\n")
    _write(tmp_path, "s.anchor", "2\n3\n3\n")
    init = api.IBStandardInitializer(2, [base])
    m, s, k, r, f = init.springs[0]
    assert list(zip(m.tolist(), s.tolist(), k.tolist(), r.tolist())) == [(0, 1, 2.0, 0.5), (2, 3, 1.0, 0.0), (0, 3, 4.0, 0.25)]
    pv, cu, nx, bend, curv = init.beams[0]
    assert (pv.tolist(), cu.tolist(), nx.tolist(), bend.tolist()) == ([0], [1], [2], [3.0]) and curv.tolist() == [[0.1, 0.2]]
    ti, tk, te = init.targets[0]
    assert ti.tolist() == [2, 1] and tk.tolist() == [10.0, 4.0] and te.tolist() == [0.5, 0.0]
    assert init.anchors[0].tolist() == [3]


@pytest.mark.parametrize("ext,text", [
    ("spring", "1\n0 9 1.0 0.0\n"),        # vertex index out of range
    ("spring", "1\n0 1 -1.0 0.0\n"),       # negative spring constant
    ("spring", "1\n0 1 1.0 -0.5\n"),       # negative rest length
    ("spring", "2\n0 1 1.0 0.0\n"),        # premature end of file
    ("spring", "0\n"),                      # invalid count
    ("beam", "1\n0 1 2 -3.0\n"),           # negative rigidity
    ("beam", "1\n0 1 2 3.0 0.5\n"),        # incomplete curvature
    ("target", "1\n1 -2.0\n"),             # negative stiffness
    ("target", "1\n7 2.0\n"),              # out of range
])
def test_invalid_entries_are_errors(tmp_path, ext, text):
    base = _write(tmp_path, "e.vertex", "3\n0 0\n1 0\n1 1\n")
    _write(tmp_path, "e." + ext, text)
    with pytest.raises(api.IBKError) as e:
        api.IBStandardInitializer(2, [base])
    assert e.value.code == api.IBK_ERR_INVALID and "e." + ext in str(e.value)


def test_missing_vertex_file_is_an_error_other_files_are_optional(tmp_path):
    with pytest.raises(api.IBKError) as e:
        api.IBStandardInitializer(2, [str(tmp_path / "nothing")])
    assert "Cannot find required vertex file" in str(e.value)
    base = _write(tmp_path, "v.vertex", "2\n0 0 0\n1 1 1\n")
    init = api.IBStandardInitializer(3, [base])
    assert init.positions().tolist() == [[0, 0, 0], [1, 1, 1]] and len(init.springs[0][0]) == 0
    with pytest.raises(api.IBKError):  # too few coordinates for NDIM = 3
        api.IBStandardInitializer(3, [_write(tmp_path, "w.vertex", "1\n0 0\n")])
