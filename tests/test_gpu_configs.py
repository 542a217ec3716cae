"""GPU parity at BASELINE.json's configurations and size-independent properties at full size.

configs[0] LakeAtRest 20 164 cells is in test_gpu_parity.py. Here:
configs[1] ClassicThacker, StructTriangMesh(512) = 1 048 576 cells, HLLC<Einfeldt>, SSPRK2
configs[2] bowl.msh refined 4x = 3 785 728 cells, paraboloid bed, HLLC<Einfeldt>, SSPRK2
configs[3] 67 108 864 cells on one GPU: mass conservation, lake at rest, symmetry (no oracle run)
plus the 10^4-step drift bound of the north star (1e-9 relative L2; we get equality).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_case, rel_l2

pytestmark = pytest.mark.gpu


def _pair(mesh, v0, reorder=False):
    from swe_fvm_b200.solver import SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, reorder=reorder)
    ref = Oracle(mesh, threads=0)  # all host threads; OpenMP oracle == scalar oracle (tested on CPU)
    ref.set_state(v0)
    return sd, TimeDisc(sd), ref


def test_config1_thacker_1m_cells():
    """1e-12 relative L2 per step is the bar (north star); the kernels give equality."""
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = make_case("classic_thacker", 512, quad_n=2)
    assert mesh.nt == 1048576
    sd, td, ref = _pair(mesh, v0)
    dt = 2e-4
    for k in range(6):
        Solvers.SSPRK2(td, dt)
        ref.step(1, 1, 2, dt)
        got, want = sd.GetVolField(), ref.get_state()
        assert rel_l2(got, want) <= 1e-12
        dt = td.CFLdt()
        assert dt == ref.cfl_dt()
    np.testing.assert_array_equal(got, want)
    cls = sd.cell_class()
    assert (cls == 1).sum() > 0 and (cls == 2).sum() > 0 and (cls == 0).sum() > 0.8 * mesh.nt  # wet/dry front present


def test_config2_bowl_refined_4x():
    from swe_fvm_b200 import TriangMesh
    from swe_fvm_b200.solver import Solvers
    m = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    for _ in range(4):
        m = m.refine()
    assert (m.nt, m.ne, m.nn) == (3785728, 5681152, 1895425)
    mesh, case, v0 = make_case("bowl_hump", mesh=m, level=3.0, amp=0.5)
    sd, td, ref = _pair(mesh, v0, reorder=True)
    m0 = sd.diagnostics()["mass"]
    dt = 1e-4
    for k in range(4):
        Solvers.SSPRK2(td, dt)
        ref.step(1, 1, 2, dt)
        dt = td.CFLdt()
        assert dt == ref.cfl_dt()
    got, want = sd.GetVolField(), ref.get_state()
    assert rel_l2(got, want) <= 1e-12
    np.testing.assert_array_equal(got, want)
    assert abs(sd.diagnostics()["mass"] - m0) <= 1e-13 * m0


def test_drift_after_1e4_steps():
    """North star: within 1e-9 after 10^4 steps with a moving wet/dry front."""
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = make_case("classic_thacker", 32, quad_n=8)
    sd, td, ref = _pair(mesh, v0)
    m0 = sd.diagnostics()["mass"]
    Solvers.run(td, "ssprk2", 10000, dt=1e-3)
    ref.run(1, 1, 2, 10000, 1e-3)
    got, want = sd.GetVolField(), ref.get_state()
    assert np.isfinite(got).all()
    assert rel_l2(got, want) <= 1e-9
    np.testing.assert_array_equal(got, want)
    assert abs(sd.diagnostics()["mass"] - m0) <= 1e-12 * m0
    assert abs(sd.time() - 10.0) < 1e-9


def test_config3_full_size_properties():
    """67M cells (the bench workload size): mass conserved to round-off over CFL-sized steps with
    a moving shoreline, depth stays non-negative, and a host round trip (swe_get_state ->
    swe_set_state) in the middle of a run changes nothing (bitwise)."""
    from swe_fvm_b200 import Case, StructTriangMesh
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    n = 4096
    mesh = StructTriangMesh(n, n, 4.0 / n)
    case = Case("classic_thacker", 2.0, 2.0, 4.0)
    case.set_bathymetry(mesh)
    v0 = case.initial_state(mesh, quad_n=1)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0)
    td = TimeDisc(sd)
    d0 = sd.diagnostics()
    Solvers.run(td, "ssprk2", 20, dt=2e-5)
    sd.synchronize()
    d1 = sd.diagnostics()
    assert abs(d1["mass"] - d0["mass"]) <= 1e-12 * d0["mass"]
    assert d1["hmin"] >= 0.0 and np.isfinite(d1["vmax"]) and d1["vmax"] > 0
    assert 0 < d1["wet_cells"] < 0.2 * mesh.nt
    straight = sd.GetVolField()
    sd.SetVolField(v0)
    Solvers.run(td, "ssprk2", 10, dt=2e-5)
    mid = sd.GetVolField()
    sd.SetVolField(mid)
    Solvers.run(td, "ssprk2", 10, dt=2e-5)
    np.testing.assert_array_equal(sd.GetVolField(), straight)
    # the dry part of the basin is untouched: w == cell bed exactly, zero velocity
    dry = sd.cell_class() == 0
    cb = mesh.centroids()[:, 2]
    assert (straight[dry, 0] == cb[dry]).all() and (straight[dry, 1:] == 0).all()


def test_lake_at_rest_full_size_strip():
    """Well-balancing at scale: a 16M-cell lake at rest over the step bump stays at rest."""
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = make_case("lake_at_rest", 2048, quad_n=1)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0)
    Solvers.run(TimeDisc(sd), "ssprk3", 50, dt=1e-4)
    d = sd.diagnostics()
    assert d["vmax"] <= 1e-13
    got = sd.GetVolField()
    assert np.abs(got[:, 0]).max() <= 1e-13


def test_device_side_cases_match_the_oracle():
    """§8 f2/f3: bathymetry + TriangAverage initial state computed on the device against the ORACLE's
    restatement of examples/Tests.h + include/PointOperations.h:20-44 (oracle.OracleCase, itself pinned bit for
    bit to upstream's TriangAverage in tests/test_ref_anchor.py), on a mesh the oracle generated itself.
    Tolerance 1e-12 absolute on O(1) fields: the device's sin/cos/exp/atan differ from glibc in the last ulp
    (floating-point tolerance, stated here); the product's HOST path must equal the oracle exactly."""
    from oracle.oracle import OracleCase, OracleStructMesh
    from swe_fvm_b200 import Case, StructTriangMesh
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    n = 96
    for kind in ("classic_thacker", "fully_wet", "lake_at_rest", "bowl_hump"):
        mesh = StructTriangMesh(n, n, 4.0 / n)
        case = Case(kind, 2.0, 2.0, 4.0)
        sd = SpaceDisc("hllc", "einfeldt", mesh, None)   # bed = 0 at creation
        sd.set_case_bathymetry(case)
        sd.set_case_state(case, quad_n=6)
        dev = sd.GetVolField()
        case.set_bathymetry(mesh)
        host = case.initial_state(mesh, quad_n=6)
        omesh = OracleStructMesh(n, n, 4.0 / n)
        ocase = OracleCase(kind, 2.0, 2.0, 4.0, level=case.c.level, amp=case.c.amp)
        ocase.set_bathymetry(omesh)
        want = ocase.initial_state(omesh, quad_n=6)
        np.testing.assert_array_equal(mesh.geometry, omesh.geometry)
        np.testing.assert_array_equal(host, want)
        assert np.abs(dev - want).max() <= 1e-12, kind
    mesh, case, v0 = make_case("classic_thacker", n, quad_n=6)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0)
    td = TimeDisc(sd)
    Solvers.run(td, "ssprk2", 100, dt=1e-3)
    t = sd.time()
    err = sd.case_l2_error(case, t)
    q, cen, A = sd.GetVolField(), mesh.centroids(), mesh.areas()
    ex = np.array([case.eval(x, y, t) for x, y in cen[:, :2]])
    h = q[:, 0] - cen[:, 2]
    want = [np.sqrt((A * (h - ex[:, 1]) ** 2).sum()), np.sqrt((A * (h * q[:, 1] - ex[:, 1] * ex[:, 2]) ** 2).sum()),
            np.sqrt((A * (h * q[:, 2] - ex[:, 1] * ex[:, 3]) ** 2).sum())]
    assert np.allclose(err, want, rtol=1e-9, atol=1e-14)
    assert err[0] < 1e-2


def test_config3_full_size_bit_parity():
    """The bench workload itself (67 108 864 cells, fully-wet variant, Hilbert numbering on the
    device): two SSPRK2 HLLC<Einfeldt> steps with dt = CFLdt on the GPU equal the (OpenMP) CPU oracle
    bit for bit, including the CFL time step."""
    from swe_fvm_b200 import Case, StructTriangMesh
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    n = 4096
    mesh = StructTriangMesh(n, n, 4.0 / n)
    case = Case("fully_wet", 2.0, 2.0, 4.0)
    case.set_bathymetry(mesh)
    v0 = case.initial_state(mesh, quad_n=1)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, reorder=True)
    td = TimeDisc(sd)
    ref = Oracle(mesh, threads=0)
    ref.set_state(v0)
    dt = 1e-5
    for _ in range(2):
        Solvers.SSPRK2(td, dt)
        ref.step(1, 1, 2, dt)
        dt = td.CFLdt()
        assert dt == ref.cfl_dt()
    got, want = sd.GetVolField(), ref.get_state()
    assert rel_l2(got, want) <= 1e-12
    np.testing.assert_array_equal(got, want)
