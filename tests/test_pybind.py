"""The pybind11 module SWE_FVM (pybind/swe_fvm_module.cpp): the working version of upstream's
pybind/Topology.cpp (whose Topology stores references to temporaries: notebooks/Untitled1.ipynb
shows num_edges() == 44070616), extended with the mesh classes and the time step."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, has_gpu, make_case


@pytest.fixture(scope="module")
def mod():
    from swe_fvm_b200 import build as b
    path = b.build_pybind()
    spec = importlib.util.spec_from_file_location("SWE_FVM", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_topology_counts_are_right(mod):
    """Same call as the reference's binding: Topology.create(numNodes, edgeNodes, edgeElements,
    elementNodes, elementEdges, elementNeighbours) (pybind/Topology.cpp:20-26)."""
    from swe_fvm_b200 import StructTriangMesh
    m = StructTriangMesh(5, 4, 0.25)
    args = [np.array(a) for a in (m.edge_nodes, m.edge_elements, m.element_nodes, m.element_edges, m.element_neighbours)]
    t = mod.Topology.create(m.nn, *args)
    del args  # the binding owns copies: no dangling references
    assert (t.num_nodes(), t.num_edges(), t.num_elements()) == (m.nn, m.ne, m.nt)
    assert t.is_edge_boundary(0)
    t2 = mod.Topology(m.nn, m.edge_nodes, m.edge_elements, m.element_nodes, m.element_edges, m.element_neighbours)
    assert t2.num_edges() == m.ne
    with pytest.raises(ValueError):
        mod.Topology.create(3, np.zeros((4, 3), np.int64), np.zeros((4, 2), np.int64), np.zeros((2, 3), np.int64),
                            np.zeros((2, 3), np.int64), np.zeros((2, 3), np.int64))


def test_mesh_classes(mod):
    from swe_fvm_b200 import TriangMesh
    g = mod.TriangMesh(os.path.join(GOLDEN, "bowl.msh"))
    ref = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    assert (g.num_nodes(), g.num_edges(), g.num_elements()) == (7555, 22342, 14788)
    np.testing.assert_array_equal(g.element_edges, ref.element_edges)
    np.testing.assert_array_equal(g.edge_elements, ref.edge_elements)
    np.testing.assert_array_equal(g.geometry[:, :2], ref.geometry[:, :2])
    g.geometry[:, 2] = -1.0  # bathymetry is writable in place
    assert (g.geometry[:, 2] == -1.0).all()
    assert not g.element_nodes.flags.writeable
    r = g.refine()
    assert r.num_elements() == 4 * g.num_elements()
    assert r.topology().num_edges() == r.num_edges()
    s = mod.StructTriangMesh(8, 6, 0.5)
    assert (s.ni(), s.nj(), s.num_elements()) == (8, 6, 192)
    with pytest.raises(RuntimeError):
        mod.TriangMesh("/nonexistent.msh")


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_space_disc_fails_loudly_without_gpu(mod):
    s = mod.StructTriangMesh(4, 4, 1.0)
    with pytest.raises(RuntimeError) as ei:
        mod.SpaceDisc("hllc", "einfeldt", s, np.zeros((s.num_elements(), 3)))
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(ValueError):
        mod.SpaceDisc("hllc", "einfeldt", s, np.zeros((3, 3)))


@pytest.mark.gpu
def test_pybind_step_equals_ctypes_path(mod):
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = make_case("classic_thacker", 48, quad_n=4)
    pm = mod.StructTriangMesh(48, 48, 4.0 / 48)
    pm.geometry[:, 2] = mesh.geometry[:, 2]
    sd = mod.SpaceDisc("hllc", "einfeldt", pm, v0, 0.1)
    td = mod.TimeDisc(sd)
    ref = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.1)
    rtd = TimeDisc(ref)
    for _ in range(5):
        mod.Solvers.ssprk2(td, 2e-3)
        Solvers.SSPRK2(rtd, 2e-3)
    mod.Solvers.run(td, "ssprk3", 5, 0.0, td.cfl_dt())
    Solvers.run(rtd, "ssprk3", 5, dt=0.0, dt0=rtd.CFLdt())
    np.testing.assert_array_equal(sd.get_vol_field(), ref.GetVolField())
    assert td.cfl_dt() == rtd.CFLdt()
    sd.compute_interface_values(); sd.compute_fluxes()
    ref.ComputeInterfaceValues(); ref.ComputeFluxes()
    np.testing.assert_array_equal(sd.get_fluxes(), ref.GetFluxes())
