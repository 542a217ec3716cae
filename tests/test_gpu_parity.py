"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): fields within 1e-12 relative L2 per step. The kernels are
written to be bit-identical with the oracle (same operation order, -fmad=false vs
-ffp-contract=off), so most checks below demand exact equality; tolerances are stated where used.
"""
import numpy as np
import pytest

from conftest import make_case, rel_l2

pytestmark = pytest.mark.gpu

SCHEMES = {"euler": 0, "ssprk2": 1, "ssprk3": 2}


def _pair(mesh, v0, flux="hllc", ws="einfeldt", cor=0.0, reorder=False, taps=True):  # noqa: D103
    from swe_fvm_b200.solver import SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    sd = SpaceDisc(flux, ws, mesh, v0, cor=cor, reorder=reorder, taps=taps)
    ref = Oracle(mesh, cor=cor)
    ref.set_state(v0)
    return sd, TimeDisc(sd), ref


def _cases():
    from swe_fvm_b200 import TriangMesh
    import os
    from conftest import GOLDEN
    out = {}
    out["lake71"] = make_case("lake_at_rest", 71)
    out["thacker64"] = make_case("classic_thacker", 64, quad_n=8)
    out["wet48"] = make_case("fully_wet", 48)
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    out["bowl_hump"] = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    return out


@pytest.fixture(scope="module")
def cases():
    return _cases()


@pytest.mark.parametrize("name", ["lake71", "thacker64", "wet48", "bowl_hump"])
@pytest.mark.parametrize("reorder", [False, True])
def test_stage_taps_bit_exact(cases, name, reorder):
    """Every intermediate of one stage equals the oracle's: classes, edge states, sources, node
    maxima, fluxes, CFL min, draining dt."""
    mesh, case, v0 = cases[name]
    sd, td, ref = _pair(mesh, v0, reorder=reorder)
    # advance a few steps first so that fronts / velocities are non-trivial
    from swe_fvm_b200.solver import Solvers
    for _ in range(3):
        Solvers.SSPRK2(td, 1e-3)
        ref.step(1, 1, 2, 1e-3)
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    sd.ComputeInterfaceValues()
    ref.compute_interface_values()
    np.testing.assert_array_equal(sd.cell_class(), ref.cell_class())
    np.testing.assert_array_equal(sd.node_max_w(), ref.node_max_w())
    np.testing.assert_array_equal(sd.GetEdgField(), ref.edge_states())
    np.testing.assert_array_equal(sd.GetSrcField(), ref.sources())
    sd.ComputeFluxes()
    ref.compute_fluxes(1, 2)
    np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes())
    assert sd.GetMinLenToWavespeed() == ref.min_len_to_wavespeed()
    assert td.CFLdt() == ref.cfl_dt()
    sd._call("swe_stage_update", 0.0, 1.0, 1e-3)
    ref.stage_update(None, 0.0, 1.0, 1e-3, True)
    np.testing.assert_array_equal(sd.draining_dt(), ref.draining_dt())
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())


@pytest.mark.parametrize("scheme", ["euler", "ssprk2", "ssprk3"])
@pytest.mark.parametrize("flux,ws", [("hllc", "einfeldt"), ("hll", "einfeldt"), ("hllc", "davis"), ("hll", "rusanov"),
                                     ("hllc", "rusanov"), ("hll", "davis")])
def test_step_parity_all_schemes_fluxes(cases, scheme, flux, ws):
    """Whole-step parity (1e-12 rel-L2 per step is the bar; we get equality) for every
    registered Fluxer x TimeDisc scheme, with Coriolis on."""
    from swe_fvm_b200 import capi
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["thacker64"]
    sd, td, ref = _pair(mesh, v0, flux=flux, ws=ws, cor=0.3, taps=False)
    step = getattr(Solvers, {"euler": "Euler", "ssprk2": "SSPRK2", "ssprk3": "SSPRK3"}[scheme])
    for k in range(20):
        step(td, 2e-3)
        ref.step(SCHEMES[scheme], capi.FLUXES[flux], capi.WAVESPEEDS[ws], 2e-3)
        got, want = sd.GetVolField(), ref.get_state()
        assert rel_l2(got, want) <= 1e-12, (k, rel_l2(got, want))
    np.testing.assert_array_equal(got, want)
    assert td.CFLdt() == ref.cfl_dt()


def _front_states(mesh, v0):
    from conftest import crafted_branch_state, random_front_state
    T = mesh.centroids()
    yield "ic", v0
    for seed, level in ((0, -0.6), (1, -0.3), (2, 0.0), (3, 7.5)):
        yield f"rough{seed}", random_front_state(mesh, T, seed, level)
    yield "crafted", crafted_branch_state(mesh, T)


def _rough_bed_case():
    from swe_fvm_b200 import StructTriangMesh
    m2 = StructTriangMesh(16, 16, 0.25)
    rng = np.random.default_rng(11)
    m2.geometry[:, 2] = 0.05 * rng.standard_normal(m2.nn)
    b13 = m2.geometry[:, 2][m2.element_nodes].max(1)
    return m2, np.stack([b13 + 10.0 ** rng.uniform(-6, -1, m2.nt), 0.1 * rng.standard_normal(m2.nt), 0.1 * rng.standard_normal(m2.nt)], 1)


@pytest.mark.parametrize("opts", [dict(), dict(recon=1, pw2=1), dict(recon=2), dict(recon=1), dict(pw2=1),
                                  dict(roe_fix=1, cfl_abs=1), dict(recon=1, pw2=1, roe_fix=1)])
@pytest.mark.parametrize("reorder", [False, True])
def test_semantic_switches_and_every_branch_bit_exact(opts, reorder):
    """swe_set_option(recon / pw2 / roe_fix / cfl_abs) against the oracle with the same switches (the oracle's
    as-written mode equals upstream's own sources bit for bit, tests/test_ref_anchor.py), on states that
    drive the reconstruction through EVERY branch of ReconstructPartWetCell1/2 and ReconstructFullWetCell;
    the device's branch-hit counters must equal the oracle's."""
    from oracle.oracle import Oracle
    from swe_fvm_b200.solver import SpaceDisc
    mesh0, case, v0 = make_case("classic_thacker", 24)
    total = np.zeros(12, dtype=np.int64)
    work = [(mesh0, n, st) for n, st in _front_states(mesh0, v0)] + [(_rough_bed_case()[0], "film", _rough_bed_case()[1])]
    for mesh, name, st in work:
        sd = SpaceDisc("hllc", "einfeldt", mesh, st, cor=0.3, reorder=reorder, taps=True)
        ref = Oracle(mesh, cor=0.3, **opts)
        ref.set_state(st)
        for k, v in opts.items():
            sd.set_option(k, v)
            assert sd.get_option(k) == v
        sd.ComputeInterfaceValues()
        ref.compute_interface_values()
        np.testing.assert_array_equal(sd.cell_class(), ref.cell_class(), err_msg=name)
        np.testing.assert_array_equal(sd.node_max_w(), ref.node_max_w(), err_msg=name)
        np.testing.assert_array_equal(sd.GetEdgField(), ref.edge_states(), err_msg=name)
        np.testing.assert_array_equal(sd.GetSrcField(), ref.sources(), err_msg=name)
        assert sd.branch_counts() == ref.branch_counts(), name
        total += np.array(list(sd.branch_counts().values()))
        for fl, ws in ((1, 2), (0, 2), (1, 1)):
            sd.flux, sd.wavespeed = fl, ws
            sd.ComputeFluxes()
            ref.compute_fluxes(fl, ws)
            np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes(), err_msg=name)
            assert sd.GetMinLenToWavespeed() == ref.min_len_to_wavespeed()
        sd.flux, sd.wavespeed = 1, 2
        for _ in range(4):  # whole steps from this state
            sd._call("swe_step", 1, 1, 2, 2e-3)
            ref.step(1, 1, 2, 2e-3)
        np.testing.assert_array_equal(sd.GetVolField(), ref.get_state(), err_msg=name)
        sd.close()
    hit = dict(zip(SpaceDisc.BRANCHES, total.tolist()))
    if opts.get("recon") == 2:  # first order: no gradient, so nothing for the vertex check / TVD test to switch off
        hit.pop("fw_vertex_zeroed"), hit.pop("fw_tvd_off")
    assert all(v > 0 for v in hit.values()), str(hit)


def test_golden_step_out1_on_the_gpu():
    """The one step for which upstream holds an output (notebooks/out1.dat: testGaussWave, Euler, HLL<Einfeldt>,
    dt = 1e-3, examples/Main.cpp:172-195) on the GPU in as-written mode: within 5 % rel-L2 of the dump (it was
    written by an intermediate revision), equal to the oracle's as-written mode bit for bit, and
    equal to upstream's own sources compiled in oracle/_ref when that library travelled with the tree."""
    import gzip
    import os
    from conftest import GOLDEN
    from oracle.oracle import Oracle
    from swe_fvm_b200 import TriangMesh
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    mesh, case, v0 = make_case("gauss_wave", mesh=bowl)
    out1 = np.loadtxt(gzip.open(os.path.join(GOLDEN, "out1.dat.gz"), "rt"))
    sd = SpaceDisc("hll", "einfeldt", mesh, v0, reorder=True)
    sd.set_option("recon", 1)
    sd.set_option("pw2", 1)
    Solvers.Euler(TimeDisc(sd), 1e-3)
    q = sd.GetVolField()
    hu, hv = q[:, 0] * q[:, 1], q[:, 0] * q[:, 2]
    assert rel_l2(hu, out1[:, 1]) < 0.05 and rel_l2(hv, out1[:, 2]) < 0.05
    assert rel_l2(q[:, 0], out1[:, 0]) < 1e-4
    o = Oracle(mesh, recon=1, pw2=1)
    o.set_state(v0)
    o.step(0, 0, 2, 1e-3)
    np.testing.assert_array_equal(q, o.get_state())
    from oracle import ref
    if os.path.exists(ref.lib_path("aswritten")):
        r = ref.Ref(mesh, variant="aswritten")
        r.set_state(v0)
        r.step(0, 0, 2, 1e-3)
        # upstream's in-place loops (S7/S8) are invisible on this fully wet flat-bed case
        np.testing.assert_array_equal(q, r.get_state())


def test_flux_registry_builtin_ids_and_a_user_flux(cases):
    """Plug-in point (i): fluxes are device functors in a compile-time registry. Selecting a built-in flux by
    registry id equals the enum path bit for bit; the sample user flux (csrc/user_fluxes.cuh, Local
    Lax-Friedrichs) is reachable by name and equals its numpy restatement on the same edge states."""
    from oracle.oracle import Oracle
    from swe_fvm_b200 import SweError
    from swe_fvm_b200.solver import SpaceDisc
    regs = SpaceDisc.registered_fluxes()
    assert regs["HLL<Rusanov>"] == 0 and regs["HLLC<Einfeldt>"] == 5 and "LocalLaxFriedrichs" in regs
    mesh, case, v0 = cases["thacker64"]
    sd = SpaceDisc("hll", "rusanov", mesh, v0, cor=0.3, taps=True)
    ref = Oracle(mesh, cor=0.3)
    ref.set_state(v0)
    sd.ComputeInterfaceValues()
    ref.compute_interface_values()
    sd.set_fluxer(regs["HLLC<Davis>"])       # overrides the ("hll", "rusanov") of the constructor
    sd.ComputeFluxes()
    ref.compute_fluxes(1, 1)
    np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes())
    assert sd.GetMinLenToWavespeed() == ref.min_len_to_wavespeed()
    sd.set_fluxer("LocalLaxFriedrichs")
    sd.ComputeFluxes()
    got, edg = sd.GetFluxes(), sd.GetEdgField()
    g = ref.geometry()
    et = mesh.edge_elements
    inner = np.nonzero(et[:, 1] >= 0)[0]
    lf, lt = et[inner, 0], et[inner, 1]
    L, R = edg[2 * inner + (lf < lt)], edg[2 * inner + (lt < lf)]
    be, n = g["E"][inner, 2], g["n0"][inner]
    hl, hr = L[:, 0] - be, R[:, 0] - be
    ul, ur = L[:, 1] * n[:, 0] + L[:, 2] * n[:, 1], R[:, 1] * n[:, 0] + R[:, 2] * n[:, 1]
    a = np.maximum(np.abs(ul) + np.sqrt(hl), np.abs(ur) + np.sqrt(hr))

    def elem(h, hu, hv):
        q = hu * n[:, 0] + hv * n[:, 1]
        wet = h > 1e-12
        hs = np.where(wet, h, 1.0)
        return np.where(wet, q, 0.0), np.where(wet, (q / hs) * hu + (0.5 * h * h) * n[:, 0], 0.0), np.where(wet, (q / hs) * hv + (0.5 * h * h) * n[:, 1], 0.0)

    Fl, Fr = elem(hl, hl * L[:, 1], hl * L[:, 2]), elem(hr, hr * R[:, 1], hr * R[:, 2])
    live = (hl + hr > 1e-10) & (a > 1e-10)
    for c, (ql, qr) in enumerate(((hl, hr), (hl * L[:, 1], hr * R[:, 1]), (hl * L[:, 2], hr * R[:, 2]))):
        want = np.where(live, 0.5 * (Fl[c] + Fr[c]) - 0.5 * a * (qr - ql), 0.0)
        np.testing.assert_allclose(got[inner, c], want, rtol=0, atol=1e-15)
    assert live.sum() > 100
    with pytest.raises(KeyError):
        sd.set_fluxer("NoSuchFlux")
    with pytest.raises(SweError):
        sd.set_fluxer(99)
    sd.set_fluxer(-1)                        # back to the constructor's enums
    sd.ComputeFluxes()
    ref.compute_fluxes(0, 0)
    np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes())


def test_host_buffer_pipeline_equals_separate_steps(cases):
    """swe_submit_step_host: a stream of independent states through upload -> one step -> download with the copies
    overlapping the compute; every result equals a plain set_state / swe_step / get_state of the same input."""
    import torch
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = cases["thacker64"]
    rng = np.random.default_rng(0)
    states = []
    for k in range(5):
        s = v0.copy()
        wet = s[:, 0] - mesh.centroids()[:, 2] > 1e-3
        s[wet, 1:] += 0.05 * rng.standard_normal((int(wet.sum()), 2))
        states.append(s)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.1, reorder=True)
    td = TimeDisc(sd)
    want = []
    for s in states:
        sd.SetVolField(s)
        Solvers.SSPRK2(td, 2e-3)
        want.append(sd.GetVolField())
    ins = [torch.from_numpy(s.copy()).pin_memory() for s in states]
    outs = [torch.empty_like(t).pin_memory() for t in ins]
    for a, b in zip(ins, outs):
        sd.submit_step_host(a.data_ptr(), b.data_ptr(), "ssprk2", 2e-3)
    sd.wait_host()
    for got, w in zip(outs, want):
        np.testing.assert_array_equal(got.numpy(), w)


@pytest.mark.parametrize("scheme", ["euler", "ssprk2", "ssprk3"])
def test_cuda_graph_replay_equals_plain_launches(cases, scheme):
    """swe_run on launch-bound meshes replays one captured CUDA graph per step (two graphs: the state buffers swap
    roles every step): identical bits, adaptive and fixed dt, odd and even step counts, mixed with single steps."""
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = cases["thacker64"]
    out = []
    for graph in (0, 1):
        sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.2, reorder=True)
        sd.set_option("graph", graph)
        td = TimeDisc(sd)
        Solvers.run(td, scheme, 7, dt=0.0, dt0=1e-3)
        Solvers.SSPRK2(td, 1e-3)                      # a plain step in between flips the buffer parity
        Solvers.run(td, scheme, 6, dt=2e-3)
        Solvers.run(td, scheme, 5, dt=0.0, dt0=0.0)   # continue adaptively with the device-resident dt
        sd.synchronize()
        out.append((sd.GetVolField(), td.CFLdt(), sd.time(), sd.launch_count()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2]
    assert out[0][3] == out[1][3]  # the graph path reports the kernels it replays


@pytest.mark.parametrize("name", ["thacker64", "bowl_hump", "lake71"])
@pytest.mark.parametrize("reorder", [False, True])
def test_fused_draining_dt_equals_separate_pass_and_oracle(cases, name, reorder):
    """The stage update computes the draining dt of its own 128-cell tile on chip (shared memory) and reads only the
    cells across tile borders from the list-driven K3': same bits as the separate k_drain pass and as the oracle,
    with and without taps, on wet/dry fronts (draining cells) and with Coriolis."""
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    mesh, case, v0 = cases[name]
    ref = Oracle(mesh, cor=0.2)
    ref.set_state(v0)
    sds = []
    for fused, taps in ((1, False), (0, False), (1, True)):
        sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.2, reorder=reorder, taps=taps)
        sd.set_option("fused_drain", fused)
        sd.set_option("graph", 0)
        assert sd.get_option("fused_drain") == fused
        sds.append((sd, TimeDisc(sd)))
    for k in range(6):
        ref.step(2, 1, 2, 2e-3)
        for sd, td in sds:
            Solvers.SSPRK3(td, 2e-3)
            np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    # the tap is only available when every cell's value was materialised
    with pytest.raises(Exception, match="fused stage update"):
        sds[0][0].draining_dt()
    sds[2][0].draining_dt()
    sds[0][0].rhs(1e-3)
    sds[0][0].draining_dt()


@pytest.mark.parametrize("name", ["thacker64", "bowl_hump", "wet48"])
@pytest.mark.parametrize("reorder", [False, True])
def test_k1_forms_are_bit_identical(cases, name, reorder):
    """The three forms of the reconstruction kernel — register-prefetched gathers (default), TMA-staged tiles,
    cp.async software pipeline — write the same edge states / gradients bit for bit and give the oracle's steps."""
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    mesh, case, v0 = cases[name]
    ref = Oracle(mesh, cor=0.1)
    ref.set_state(v0)
    sds = []
    for mode in (0, 1, 2):
        sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.1, reorder=reorder, taps=True)
        sd.set_option("k1_tiled", mode)
        sd.set_option("graph", 0)
        assert sd.get_option("k1_tiled") == mode
        sds.append((sd, TimeDisc(sd)))
    for k in range(5):
        ref.step(1, 1, 2, 2e-3)
        for sd, td in sds:
            Solvers.SSPRK2(td, 2e-3)
            np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    ref.compute_interface_values()
    for sd, td in sds:
        sd.ComputeInterfaceValues()
        np.testing.assert_array_equal(sd.GetEdgField(), ref.edge_states())
        np.testing.assert_array_equal(sd.GetSrcField(), ref.sources())
        np.testing.assert_array_equal(sd.cell_class(), ref.cell_class())


def _delaunay_mesh(seed, npts):
    """Random Delaunay triangulation of a jittered disc: node degrees 3..10+, ragged boundary, slivers, rough bed."""
    from scipy.spatial import Delaunay
    from swe_fvm_b200 import TriangMesh
    rng = np.random.default_rng(seed)
    r, th = np.sqrt(rng.uniform(0, 1, npts)) * 3.0, rng.uniform(0, 2 * np.pi, npts)
    xy = np.stack([4 + r * np.cos(th), 4 + r * np.sin(th)], 1)
    tri = Delaunay(xy).simplices.astype(np.int64)
    a, b, c = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
    det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tri = tri[np.abs(det) > 1e-9]  # drop degenerate slivers of the hull
    m = TriangMesh.from_triangles(xy, tri)
    g = np.asarray(m.geometry)
    g[:, 2] = 0.25 * ((g[:, 0] - 4) ** 2 + (g[:, 1] - 4) ** 2) / 9.0 + 0.03 * rng.standard_normal(m.nn)
    return m


@pytest.mark.parametrize("seed,npts", [(1, 400), (2, 3000), (3, 12000)])
@pytest.mark.parametrize("reorder", [False, True])
def test_random_delaunay_meshes_bit_exact(seed, npts, reorder):
    """Irregular topology the structured and Gmsh fixtures do not have (node degree 3..10+, boundary triangles with two
    wall edges, slivers) with rough beds and wet/dry fronts: every stage tap and whole SSPRK3 steps equal the oracle."""
    from conftest import random_front_state
    from swe_fvm_b200.solver import Solvers
    mesh = _delaunay_mesh(seed, npts)
    T = mesh.centroids()
    seen = set()
    for k, level in enumerate((0.05, 0.12, 0.4)):
        v0 = random_front_state(mesh, T, seed + k, level, amp=0.05)
        sd, td, ref = _pair(mesh, v0, cor=0.1, reorder=reorder)
        sd.ComputeInterfaceValues(); ref.compute_interface_values()
        np.testing.assert_array_equal(sd.cell_class(), ref.cell_class())
        np.testing.assert_array_equal(sd.node_max_w(), ref.node_max_w())
        np.testing.assert_array_equal(sd.GetEdgField(), ref.edge_states())
        np.testing.assert_array_equal(sd.GetSrcField(), ref.sources())
        sd.ComputeFluxes(); ref.compute_fluxes(1, 2)
        np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes())
        assert sd.GetMinLenToWavespeed() == ref.min_len_to_wavespeed()
        dt = 0.5 * td.CFLdt()
        for _ in range(4):
            Solvers.SSPRK3(td, dt)
            ref.step(2, 1, 2, dt)
            np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
        seen |= set(np.unique(ref.cell_class()).tolist())
    assert seen == {0, 1, 2}, "the cases must cover dry, part-wet and full-wet cells"


@pytest.mark.parametrize("scheme", ["euler", "ssprk2", "ssprk3"])
@pytest.mark.parametrize("reorder", [False, True])
def test_dry_region_skipping_is_bit_identical(scheme, reorder):
    """Tiles of 128 deep-dry cells (dry, stored as (cb, +0, +0), all neighbours dry) are skipped by the flux / draining-dt /
    update kernels. Same bits as doing the work and as the oracle while the shoreline of a Thacker basin sweeps over the
    tiles (wetting and drying), from a start state whose dry cells are NOT canonical (|h| <= 1e-12, -0.0, non-zero
    velocities), through swe_run (graph replay and plain launches) and single steps."""
    from swe_fvm_b200 import capi
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    from oracle.oracle import Oracle
    mesh, case, v0 = make_case("classic_thacker", 96, quad_n=4)
    v0 = v0.copy()
    T = mesh.centroids()
    dry = (v0[:, 0] - T[:, 2]) <= 1e-12
    rng = np.random.default_rng(5)
    k = np.nonzero(dry)[0]
    v0[k[::3], 0] += 7e-13                    # dry but not canonical: the first update must still canonicalise them
    v0[k[1::3], 0] -= 3e-13
    v0[k[::5], 1] = -0.0
    v0[k[2::7], 2] = 0.25 * rng.standard_normal(len(k[2::7]))
    ref = Oracle(mesh, cor=0.05)
    ref.set_state(v0)
    runs = []
    for skip, graph, dlist in ((1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)):
        sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.05, reorder=reorder)
        sd.set_option("dry_skip", skip)
        sd.set_option("graph", graph)
        sd.set_option("dry_list", dlist)  # 1: the stage update runs over the compacted list of tiles (the form big meshes use)
        assert sd.get_option("dry_skip") == skip
        runs.append((sd, TimeDisc(sd)))
    sc = SCHEMES[scheme]
    dt = 4e-3
    for block in range(6):
        for _ in range(25):
            ref.step(sc, 1, 2, dt)
        for sd, td in runs:
            Solvers.run(td, scheme, 24, dt=dt)
            getattr(Solvers, {"euler": "Euler", "ssprk2": "SSPRK2", "ssprk3": "SSPRK3"}[scheme])(td, dt)
            got = sd.GetVolField()
            np.testing.assert_array_equal(got.view(np.uint64), ref.get_state().view(np.uint64))
            assert td.CFLdt() == ref.cfl_dt()
    # fluxes and residuals of the next stage: they depend on edge states the reconstruction did not rewrite for tiles
    # that stayed deep dry (its stores are skipped while a tile stays flagged)
    ref.compute_interface_values(); ref.compute_fluxes(1, 2)
    for sd, td in runs:
        sd.ComputeInterfaceValues(); sd.ComputeFluxes()
        np.testing.assert_array_equal(sd.GetFluxes(), ref.fluxes())
        assert sd.GetMinLenToWavespeed() == ref.min_len_to_wavespeed()
        np.testing.assert_array_equal(sd.rhs(dt)[::7], np.array([ref.rhs(i, dt) for i in range(0, mesh.nt, 7)]))
    # the shoreline moved: some cells changed between dry and wet during the run, and most tiles were skipped
    ref.compute_interface_values()
    cls = ref.cell_class()
    assert (cls == 0).mean() > 0.5 and (cls == 2).any() and (cls == 1).any()
    assert ((cls == 0) != dry).any()


def test_create_rejects_another_local_edge_order():
    """swe_create validates the local convention the kernels rely on (edge k joins nodes k, k+1; neighbour k across it)."""
    from swe_fvm_b200 import StructTriangMesh, SweError
    from swe_fvm_b200.solver import SpaceDisc
    mesh = StructTriangMesh(4, 4, 1.0)
    good_e, good_t = mesh.element_edges.copy(), mesh.element_neighbours.copy()
    mesh.element_edges[:] = np.roll(good_e, 1, axis=1)      # e.g. "edge k opposite node k"
    mesh.element_neighbours[:] = np.roll(good_t, 1, axis=1)
    with pytest.raises(SweError) as ei:
        SpaceDisc("hllc", "einfeldt", mesh)
    assert ei.value.status == -1 and "element_edges[k] must join" in str(ei.value)
    mesh.element_edges[:] = good_e                           # edges right, neighbours rotated
    with pytest.raises(SweError) as ei:
        SpaceDisc("hllc", "einfeldt", mesh)
    assert "element_neighbours[k] must be the cell across" in str(ei.value)


def test_lake_at_rest_config0(cases):
    """configs[0]: LakeAtRest on StructTriangMesh(71,71,4/71), HLLC<Einfeldt>, Euler, dt=1e-3:
    velocities stay at machine zero, w stays 0."""
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["lake71"]
    sd, td, ref = _pair(mesh, v0, taps=False)
    assert mesh.nt == 20164
    for _ in range(200):
        Solvers.Euler(td, 1e-3)
    ref.run(0, 1, 2, 200, 1e-3)
    got = sd.GetVolField()
    np.testing.assert_array_equal(got, ref.get_state())
    # machine zero: a few ulp of the O(1) depths (n = 71 does not align the bump with grid
    # lines, so some cells have a sloping bed; the GPU equals the oracle bit for bit above)
    assert np.abs(got[:, 1:]).max() <= 1e-14
    assert np.abs(got[:, 0]).max() <= 1e-14


def test_adaptive_run_matches_oracle(cases):
    """swe_run with dt = CFLdt() of the previous step, all on device, vs the oracle's loop."""
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["thacker64"]
    sd, td, ref = _pair(mesh, v0, taps=False)
    Solvers.run(td, "ssprk2", 50, dt=0.0, dt0=1e-3)
    ref.run(1, 1, 2, 50, 0.0, 1e-3)
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    assert td.CFLdt() == ref.cfl_dt()
    # mass conserved to round-off (Jacobi semantics, S7)
    d = sd.diagnostics()
    o = ref.diagnostics()
    assert abs(d["mass"] - o[0]) <= 1e-13 * abs(o[0])


def test_mass_conservation_and_diagnostics(cases):
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["bowl_hump"]
    sd, td, ref = _pair(mesh, v0, taps=False)
    m0 = sd.diagnostics()["mass"]
    Solvers.run(td, "ssprk2", 100, dt=1e-3)
    d = sd.diagnostics()
    assert abs(d["mass"] - m0) <= 1e-13 * abs(m0)
    ref.run(1, 1, 2, 100, 1e-3)
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    o = ref.diagnostics()
    assert d["wet_cells"] == int(o[5])
    for k, key in enumerate(["mass", "kinetic", "potential"]):
        assert abs(d[key] - o[k]) <= 1e-12 * max(abs(o[k]), 1e-30)
    assert d["vmax"] == o[3] and d["hmin"] == o[4]


def test_checkpoint_restart_is_exact(cases, tmp_path):
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["thacker64"]
    sd, td, ref = _pair(mesh, v0, taps=False)
    Solvers.run(td, "ssprk2", 40, dt=2e-3)
    straight = sd.GetVolField()
    sd.SetVolField(v0)
    Solvers.run(td, "ssprk2", 15, dt=2e-3)
    sd.save_checkpoint(str(tmp_path / "ck.bin"))
    sd2, td2, _ = _pair(mesh, v0, reorder=True, taps=False)
    t = sd2.load_checkpoint(str(tmp_path / "ck.bin"))
    assert abs(t - 15 * 2e-3) < 1e-15
    Solvers.run(td2, "ssprk2", 25, dt=2e-3)
    np.testing.assert_array_equal(sd2.GetVolField(), straight)
    assert abs(sd2.time() - 40 * 2e-3) < 1e-14


def test_adaptive_restart_is_bit_identical(cases, tmp_path):
    """swe_checkpoint_save / _load (C-ABI) keep time, the device-resident dt and min_len: an adaptive run that
    is interrupted, written to disk and continued in ANOTHER context (other device numbering) equals the
    uninterrupted run bit for bit; a checkpoint of another mesh / other settings is refused."""
    from swe_fvm_b200 import SweError
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["thacker64"]
    sd, td, _ = _pair(mesh, v0, taps=False, cor=0.2)
    Solvers.run(td, "ssprk2", 40, dt=0.0, dt0=1e-3)
    straight, t_straight = sd.GetVolField(), sd.time()
    sd.SetVolField(v0)
    Solvers.run(td, "ssprk2", 17, dt=0.0, dt0=1e-3)
    sd.save_checkpoint(str(tmp_path / "ck.bin"))
    sd2, td2, _ = _pair(mesh, v0, reorder=True, taps=False, cor=0.2)
    sd2.load_checkpoint(str(tmp_path / "ck.bin"))
    Solvers.run(td2, "ssprk2", 23, dt=0.0, dt0=0.0)   # dt0 <= 0: continue with the restored dt
    np.testing.assert_array_equal(sd2.GetVolField(), straight)
    assert sd2.time() == t_straight
    sd3, _, _ = _pair(mesh, v0, taps=False, cor=0.0)      # other Coriolis parameter
    with pytest.raises(SweError):
        sd3.load_checkpoint(str(tmp_path / "ck.bin"))
    other, _, v1 = cases["wet48"]
    sd4, _, _ = _pair(other, v1, taps=False, cor=0.2)     # other mesh
    with pytest.raises(SweError):
        sd4.load_checkpoint(str(tmp_path / "ck.bin"))
    with pytest.raises(SweError):
        sd4.load_checkpoint(str(tmp_path / "missing.bin"))


def test_reference_style_time_loop_with_rhs_and_cell_classes(cases):
    """The reference's own loop shape, `cons(i) += td->RHS(i, dt)` (src/Solvers.cpp:11-13), expressed with the
    per-cell accessors: RHS(i, dt) and Is*Cell(i) are whole-array taps of the device. One Euler step assembled
    on the host from sd.rhs(dt) through the oracle's ConsAssigner equals swe_step bit for bit."""
    from oracle.oracle import Oracle
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = cases["thacker64"]
    sd, td, ref = _pair(mesh, v0, cor=0.3)
    for _ in range(3):
        Solvers.SSPRK2(td, 2e-3)
        ref.step(1, 1, 2, 2e-3)
    np.testing.assert_array_equal(sd.classify(), ref.cell_class())
    i_dry, i_pw, i_fw = (int(np.nonzero(ref.cell_class() == k)[0][0]) for k in (0, 1, 2))
    assert sd.IsDryCell(i_dry) and sd.IsPartWetCell(i_pw) and sd.IsFullWetCell(i_fw)
    sd.ComputeInterfaceValues(); sd.ComputeFluxes()
    ref.compute_interface_values(); ref.compute_fluxes(1, 2)
    dt = 3e-3
    rhs = sd.rhs(dt)
    want = np.array([ref.rhs(i, dt) for i in range(0, mesh.nt, 37)])
    np.testing.assert_array_equal(rhs[::37], want)
    np.testing.assert_array_equal(sd.draining_dt(), ref.draining_dt_live())
    assert td.RHS(i_fw, dt).tolist() == rhs[i_fw].tolist() and td.ComputeDrainingDt(-1) == float("inf")
    # host-side Euler step from the taps (ConsAssigner += RHS), then the device's own step
    state = sd.GetVolField()
    o2 = Oracle(mesh, cor=0.3)
    o2.set_state(state)
    for i in range(mesh.nt):
        o2.assign_cons(i, o2.get_cons(i) + rhs[i])
    Solvers.Euler(td, dt)
    np.testing.assert_array_equal(sd.GetVolField(), o2.get_state())


@pytest.mark.parametrize("ni,nj", [(1, 1), (2, 1), (3, 3)])
def test_tiny_meshes_all_boundary(ni, nj):
    """Edge cases: meshes where every (or almost every) triangle touches the wall, random wet / dry /
    nearly dry states, with Coriolis: GPU == oracle."""
    from swe_fvm_b200 import StructTriangMesh
    from swe_fvm_b200.solver import Solvers
    mesh = StructTriangMesh(ni, nj, 0.7)
    rng = np.random.default_rng(ni * 10 + nj)
    mesh.geometry[:, 2] = rng.uniform(-1.0, 0.2, mesh.nn)
    cb = mesh.centroids()[:, 2]
    v0 = np.zeros((mesh.nt, 3))
    depth = rng.choice([0.0, 1e-13, 5e-4, 0.05, 0.6], mesh.nt)
    v0[:, 0] = cb + depth
    v0[:, 1:] = rng.uniform(-0.3, 0.3, (mesh.nt, 2)) * (depth[:, None] > 1e-12)
    sd, td, ref = _pair(mesh, v0, cor=0.5)
    for _ in range(25):
        Solvers.SSPRK3(td, 5e-3)
        ref.step(2, 1, 2, 5e-3)
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    assert np.isfinite(sd.GetVolField()).all()


def test_all_dry_and_zero_steps(cases):
    from swe_fvm_b200.solver import Solvers
    mesh, case, v0 = cases["thacker64"]
    dry = v0.copy()
    dry[:, 0] = mesh.centroids()[:, 2]
    dry[:, 1:] = 0.0
    sd, td, ref = _pair(mesh, dry, taps=False)
    Solvers.run(td, "ssprk2", 0, dt=1e-3)       # zero steps: nothing happens
    np.testing.assert_array_equal(sd.GetVolField(), dry)
    Solvers.run(td, "ssprk2", 3, dt=1e-3)
    ref.run(1, 1, 2, 3, 1e-3)
    np.testing.assert_array_equal(sd.GetVolField(), ref.get_state())
    np.testing.assert_array_equal(sd.GetVolField(), dry)  # a dry basin stays dry
    assert td.CFLdt() == 0.15  # no wet edge: min_len keeps its reset value 1.0


def test_non_finite_state_is_reported(cases):
    """The device raises a flag when an update produces a non-finite state; swe_synchronize turns
    it into SWE_ERR_NUMERIC (SolverError in the C++ shim) instead of silently continuing."""
    from swe_fvm_b200 import SweError
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = cases["wet48"]
    bad = v0.copy()
    bad[mesh.nt // 2, 1] = np.nan
    sd = SpaceDisc("hllc", "einfeldt", mesh, bad)
    Solvers.SSPRK2(TimeDisc(sd), 1e-3)
    with pytest.raises(SweError) as ei:
        sd.synchronize()
    assert ei.value.status == -4


def test_argument_errors(cases):
    from swe_fvm_b200 import SweError
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    mesh, case, v0 = cases["wet48"]
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0)
    td = TimeDisc(sd)
    with pytest.raises(SweError):
        Solvers.SSPRK2(td, -1.0)
    with pytest.raises(SweError):
        sd._call("swe_step", 7, 1, 2, 1e-3)
    with pytest.raises(SweError):
        sd._call("swe_compute_fluxes", 5, 2)
    with pytest.raises(SweError):
        sd.GetEdgField()  # taps were not enabled
    with pytest.raises(ValueError):
        sd.SetVolField(v0[:-1])
