"""Mesh2D.refine and Mesh2D.embedding (learnmultigrid_b200/mesh/Mesh2D.py) against the reference's own methods
(tests/golden/mesh_embedding.npz, made by tests/golden/make_golden_mesh.py through oracle/refshim.py):
node coordinates bit for bit, element lists exact."""
import numpy as np
import pytest

from helpers import load_golden

EMBED_CASES = ["s4", "s12", "s16", "s60", "s64", "r4", "i16x2", "i12"]
REFINE_CASES = [(4, True, 0, 2), (16, False, 42, 2), (12, False, 3, 1), (60, False, 1, 1)]


@pytest.mark.parametrize("name", EMBED_CASES)
def test_embedding_equals_reference(name):
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    d = load_golden("mesh_embedding.npz")
    mesh = Mesh2D(p=d[name + "_in_p"], conn=d[name + "_in_conn"].astype(int))
    before = (mesh.p.copy(), mesh.conn.copy())
    e = mesh.embedding()
    assert np.array_equal(e.p, d[name + "_p"])
    assert np.array_equal(e.conn, d[name + "_conn"])
    assert e.n_p == len(d[name + "_p"]) and e.ne == len(d[name + "_conn"])
    assert np.array_equal(mesh.p, before[0]) and np.array_equal(mesh.conn, before[1])     # input mesh untouched
    # the original nodes and elements keep their numbers; two ghost layers: (W+4)(H+4) nodes in all
    assert np.array_equal(e.p[:mesh.n_p], mesh.p) and np.array_equal(e.conn[:mesh.ne], mesh.conn)


@pytest.mark.parametrize("ne,regular,seed,times", REFINE_CASES)
def test_refine_equals_reference(ne, regular, seed, times):
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    d = load_golden("mesh_embedding.npz")
    name = "ref_%d_%d_%d_%d" % (ne, int(regular), seed, times)
    np.random.seed(seed)
    m = Mesh2D(ne)
    for _ in range(times):
        m.refine(regular=regular)
    assert np.array_equal(m.p, d[name + "_p"])
    assert np.array_equal(m.conn, d[name + "_conn"])
    assert m.n_p == len(m.p) and m.ne == len(m.conn)


def test_embedding_of_structured_mesh_is_the_extended_lattice():
    """size-independent property: the embedding of an (n x n)-square mesh is the lattice with two more squares on
    every side, every ghost square split into two triangles of positive area"""
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    n = 32
    e = Mesh2D(n * n).embedding()
    assert e.n_p == (n + 5) ** 2 and e.ne == 2 * (n + 4) ** 2
    h = 1.0 / n
    lattice = {(round(x / h), round(y / h)) for x, y in e.p}
    assert lattice == {(i, j) for i in range(-2, n + 3) for j in range(-2, n + 3)}
    a, b, c = (e.p[e.conn[:, k]] for k in range(3))
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    np.testing.assert_allclose(area, 0.5 * h * h, rtol=1e-12)
    np.testing.assert_allclose(area.sum(), (1 + 4 * h) ** 2, rtol=1e-12)


def test_embedded_mass_matrix_fills_boundary_patches():
    """what the embedding is for (test/test_mg_2d_with_embedding.py:286-316): on the embedded mesh every original
    node, boundary nodes included, has a full interior mass-matrix row"""
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    from learnmultigrid_b200.assembly.MassMatrix import MassMatrix
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import FunctionTriangle
    mesh = Mesh2D(64)
    M = MassMatrix(mesh.embedding()).compute_mass_2d(FunctionTriangle(1), Quadrature2D(3), format="csr")
    rows = M[:mesh.n_p]
    assert np.all(np.diff(rows.indptr) == 7)
    h2 = (1.0 / 8) ** 2
    np.testing.assert_allclose(rows.diagonal(), h2 / 2, rtol=1e-12)
    np.testing.assert_allclose(np.asarray(rows.sum(axis=1)).ravel(), h2, rtol=1e-12)


@pytest.mark.parametrize("ne,key", [(16, "N4"), (256, "N16"), (12, "rect12")])
def test_structured_mesh_construction_equals_reference(ne, key):
    """Mesh2D(ne): node coordinates and element list of the reference (Mesh2D.py:41-93), including the rectangular
    4 x 3 split that find_balanced_couple picks for ne = 12 (tests/golden/assembly_2d.npz)"""
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    d = load_golden("assembly_2d.npz")
    m = Mesh2D(ne)
    assert np.array_equal(m.get_points(), d[key + "_p"])
    assert np.array_equal(m.get_connections(), d[key + "_conn"])
    assert m.get_np() == len(d[key + "_p"]) and m.get_ne() == len(d[key + "_conn"])
