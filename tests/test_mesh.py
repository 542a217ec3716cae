"""Host mesh builders against the reference's own fixtures (no GPU needed)."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _load_topology():
    with gzip.open(os.path.join(GOLDEN, "topology.dat.gz"), "rt") as f:
        lines = f.read().split("\n")
    ne = int(lines[0])
    E = np.array([l.split() for l in lines[1:1 + ne]], dtype=np.int64)
    nt = int(lines[1 + ne])
    T = np.array([l.split() for l in lines[2 + ne:2 + ne + nt]], dtype=np.int64)
    return E, T


def test_gmsh_reader_reproduces_reference_numbering():
    """notebooks/topology.dat is the reference's own dump (examples/Main.cpp:146-165) of
    examples/bowl.msh: every edge and triangle row must match exactly."""
    from swe_fvm_b200 import TriangMesh
    m = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    E, T = _load_topology()
    assert (m.nn, m.ne, m.nt) == (7555, 22342, 14788)
    np.testing.assert_array_equal(m.edge_nodes, E[:, :2])
    np.testing.assert_array_equal(m.edge_elements, E[:, 2:])
    np.testing.assert_array_equal(m.element_nodes, T[:, :3])
    np.testing.assert_array_equal(m.element_edges, T[:, 3:6])
    np.testing.assert_array_equal(m.element_neighbours, T[:, 6:])


def test_gmsh_reader_geometry_matches_dump():
    from swe_fvm_b200 import TriangMesh
    m = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    g = np.loadtxt(gzip.open(os.path.join(GOLDEN, "geometry.dat.gz"), "rt"))
    assert g.shape == (m.nn, 3)
    assert np.abs(g[:, :2] - m.geometry[:, :2]).max() <= 5e-6 * 8  # 6 significant digits on [0, 8]


def test_gmsh_42_with_physical_names():
    """notebooks/basic.msh: format 4.2, $PhysicalNames and a trailing $Projection section."""
    from swe_fvm_b200 import TriangMesh
    m = TriangMesh.from_gmsh(os.path.join(GOLDEN, "basic.msh"))
    assert m.nn == 160 and m.nt == 258
    assert m.ne == m.nn + m.nt - 1  # Euler characteristic of a disc
    assert (m.edge_elements[:, 1] < 0).sum() == 60  # boundary line elements


def test_gmsh_reader_errors():
    from swe_fvm_b200 import SweError, TriangMesh
    with pytest.raises(SweError):
        TriangMesh.from_gmsh("/nonexistent/file.msh")


@pytest.mark.parametrize("ni,nj", [(1, 1), (3, 2), (7, 5), (16, 16)])
def test_struct_mesh_follows_generic_numbering_rules(ni, nj):
    """The closed-form StructTriangMesh generator equals the generic first-visit builder applied
    to its own triangles, so one set of conventions serves both mesh sources."""
    from swe_fvm_b200 import StructTriangMesh, TriangMesh
    m = StructTriangMesh(ni, nj, 0.25)
    g = TriangMesh.from_triangles(m.geometry[:, :2].copy(), m.element_nodes.copy())
    for name in ("edge_nodes", "edge_elements", "element_nodes", "element_edges", "element_neighbours"):
        np.testing.assert_array_equal(getattr(m, name), getattr(g, name), err_msg=name)
    assert m.nt == 4 * ni * nj and m.ne == 6 * ni * nj + ni + nj and m.nn == (ni + 1) * (nj + 1) + ni * nj


def _check_conventions(m):
    tp, te, tt, ep, et = m.element_nodes, m.element_edges, m.element_neighbours, m.edge_nodes, m.edge_elements
    xy = m.geometry[:, :2]
    a, b, c = xy[tp[:, 0]], xy[tp[:, 1]], xy[tp[:, 2]]
    det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    assert (det > 0).all()  # CCW
    assert (ep[:, 0] < ep[:, 1]).all()  # sorted edge nodes
    for k in range(3):
        e = te[:, k]
        lo = np.minimum(tp[:, k], tp[:, (k + 1) % 3])
        hi = np.maximum(tp[:, k], tp[:, (k + 1) % 3])
        assert (ep[e, 0] == lo).all() and (ep[e, 1] == hi).all()  # edge k joins ip[k], ip[k+1]
        i = np.arange(m.nt)
        other = np.where(et[e, 0] == i, et[e, 1], et[e, 0])
        assert (other == tt[:, k]).all()  # neighbour k is across edge k
        assert ((et[e, 0] == i) | (et[e, 1] == i)).all()
    interior = et[:, 1] >= 0
    assert (et[interior, 0] > et[interior, 1]).all()  # (later, earlier)
    assert (et[~interior, 1] == -1).all()  # SOLID_WALL


def test_conventions_struct_and_refined():
    from swe_fvm_b200 import StructTriangMesh, TriangMesh
    _check_conventions(StructTriangMesh(9, 6, 0.5))
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    _check_conventions(bowl)
    r = bowl.refine()
    assert (r.nt, r.ne, r.nn) == (4 * bowl.nt, 2 * bowl.ne + 3 * bowl.nt, bowl.nn + bowl.ne)
    _check_conventions(r)
    assert abs(r.areas().sum() - bowl.areas().sum()) <= 1e-10 * bowl.areas().sum()


def test_struct_block_offsets_are_bitwise_global():
    """A strip built with (i0, j0) offsets has exactly the node coordinates of the global mesh."""
    from swe_fvm_b200 import StructTriangMesh
    n, h = 12, 4.0 / 12
    g = StructTriangMesh(n, n, h)
    s = StructTriangMesh(n, 5, h, 0, 4)
    gc = g.centroids()[4 * n * 4: 4 * n * 9]
    np.testing.assert_array_equal(s.centroids()[:, :2], gc[:, :2])


def test_partition_and_extract():
    from swe_fvm_b200 import TriangMesh
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    part = bowl.partition_rcb(3)
    counts = np.bincount(part, minlength=3)
    assert counts.sum() == bowl.nt and counts.max() - counts.min() <= 2
    for r in range(3):
        sub = bowl.extract(part, r, 2)
        gc = np.array(sub.global_cells)
        assert (np.diff(gc) > 0).all()  # increasing global id => orientation preserved
        assert set(np.nonzero(part == r)[0]) <= set(gc)
        _check_conventions_sub(sub, bowl, gc)


def _check_conventions_sub(sub, g, gc):
    # local triangles are the global ones (same node coordinates, same local order)
    np.testing.assert_array_equal(sub.geometry[sub.element_nodes][:, :, :2], g.geometry[g.element_nodes[gc]][:, :, :2])
    # interior edges keep the global (from, to) orientation
    et = sub.edge_elements
    interior = et[:, 1] >= 0
    assert (gc[et[interior, 0]] > gc[et[interior, 1]]).all()


def test_cases_match_reference_formulas():
    """examples/Tests.h: LakeAtRest bed and ClassicThacker at t = 0 with the suggested defaults."""
    from swe_fvm_b200 import Case
    lake = Case("lake_at_rest", 2, 2, 4)
    assert lake.eval(2.0, 2.0)[0] == -0.2 and lake.eval(0.5, 2.0)[0] == -1.0 and lake.eval(1.0, 2.0)[0] == -1.0
    th = Case("classic_thacker", 2, 2, 4)
    b, h, u, v = th.eval(2.0, 2.0, 0.0)
    assert b == -1.0 and abs(h - 0.5) < 1e-15 and u == 0 and v == 0
    # omega = sqrt(8), a = 0.6: shoreline radius at t = 0 is sqrt(2*Hc/|Hxx|) = 0.3535...
    r = 0.5 * np.sqrt(0.5)
    assert th.eval(2.0 + r * 0.999, 2.0)[1] > 0 and th.eval(2.0 + r * 1.001, 2.0)[1] == 0
    # half a period later the centre depth is H0 (1-a)/(1+a) = 0.125
    assert abs(th.eval(2.0, 2.0, np.pi / np.sqrt(8.0))[1] - 0.125) < 1e-14
