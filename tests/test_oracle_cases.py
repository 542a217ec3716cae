"""The oracle's own mesh generator, analytic cases and cell-average initial conditions (oracle/swe_oracle.cpp,
restated from SURVEY App. B, examples/Tests.h and include/PointOperations.h:20-44) against the product's HOST
code (csrc/hostmesh.cpp): two independent statements must agree exactly. No GPU needed."""
import numpy as np
import pytest

from oracle.oracle import OracleCase, OracleStructMesh


@pytest.mark.parametrize("ni,nj,i0,j0", [(1, 1, 0, 0), (5, 3, 0, 0), (7, 9, 2, 5), (32, 32, 0, 0)])
def test_struct_mesh_generators_agree(ni, nj, i0, j0):
    from swe_fvm_b200 import StructTriangMesh
    a, b = OracleStructMesh(ni, nj, 0.3, i0, j0), StructTriangMesh(ni, nj, 0.3, i0, j0)
    assert (a.nn, a.ne, a.nt) == (b.nn, b.ne, b.nt)
    for k in ("geometry", "edge_nodes", "edge_elements", "element_nodes", "element_edges", "element_neighbours"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)
    # App. B rules, checked directly: CCW, edge k joins nodes k and k+1, sorted edge nodes, later visitor first
    p = a.geometry[a.element_nodes]
    det = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1])
    assert (det > 0).all()
    for k in range(3):
        e = a.element_edges[:, k]
        pair = np.sort(np.stack([a.element_nodes[:, k], a.element_nodes[:, (k + 1) % 3]], 1), 1)
        np.testing.assert_array_equal(a.edge_nodes[e], pair)
    inner = a.edge_elements[:, 1] >= 0
    assert (a.edge_elements[inner, 0] > a.edge_elements[inner, 1]).all()


@pytest.mark.parametrize("kind,kw", [("lake_at_rest", {}), ("classic_thacker", {}), ("classic_thacker", dict(cor=0.3, q0=0.4, p0=0.1)),
                                     ("fully_wet", {}), ("bowl_hump", dict(level=0.3, amp=0.5)), ("gauss_wave", {})])
def test_cases_and_initial_states_agree(kind, kw):
    from swe_fvm_b200 import Case, StructTriangMesh
    m1, m2 = OracleStructMesh(12, 12, 4 / 12), StructTriangMesh(12, 12, 4 / 12)
    oc, pc = OracleCase(kind, 2.0, 2.0, 4.0, **kw), Case(kind, 2.0, 2.0, 4.0, **kw)
    oc.set_bathymetry(m1)
    pc.set_bathymetry(m2)
    np.testing.assert_array_equal(m1.geometry, m2.geometry)
    rng = np.random.default_rng(0)
    for x, y, t in rng.uniform(0, 4, (50, 3)):
        np.testing.assert_array_equal(oc.eval(x, y, 0.2 * t), pc.eval(x, y, 0.2 * t))
    for q in (1, 4, 7):
        t = 0.1 if kind == "classic_thacker" else 0.0
        np.testing.assert_array_equal(oc.initial_state(m1, q, t), pc.initial_state(m2, q, t))


def test_thacker_is_an_exact_solution_of_the_g1_system():
    """Finite-difference residual of the restated ClassicThackerTest (examples/Tests.h:237-280) in the g = 1
    shallow-water equations over the paraboloid bed — the formulas are a solution, not just a transcription."""
    oc = OracleCase("classic_thacker", 2.0, 2.0, 4.0, cor=0.0, q0=0.5, H0=0.5)
    e = 1e-5
    worst = 0.0
    for x, y, t in [(2.1, 2.05, 0.1), (1.9, 2.2, 0.3), (2.2, 1.8, 0.5)]:
        f = lambda a, b, c: oc.eval(a, b, c)  # noqa: E731  (b, h, u, v)
        b0, h0, u0, v0 = f(x, y, t)
        d = lambda i, axis: ((f(x + e, y, t)[i] - f(x - e, y, t)[i]) / (2 * e) if axis == 0 else  # noqa: E731
                             (f(x, y + e, t)[i] - f(x, y - e, t)[i]) / (2 * e) if axis == 1 else
                             (f(x, y, t + e)[i] - f(x, y, t - e)[i]) / (2 * e))
        mass = d(1, 2) + d(1, 0) * u0 + h0 * d(2, 0) + d(1, 1) * v0 + h0 * d(3, 1)
        momx = d(2, 2) + u0 * d(2, 0) + v0 * d(2, 1) + d(1, 0) + d(0, 0)
        momy = d(3, 2) + u0 * d(3, 0) + v0 * d(3, 1) + d(1, 1) + d(0, 1)
        worst = max(worst, abs(mass), abs(momx), abs(momy))
    assert worst < 1e-8
