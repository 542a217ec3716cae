"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo jobs run the product's C++ decomposition
plans (csrc/distplan.cpp) with a test-side exchange / dt min-reduction loop and the oracle as the local solver. The owned cells of all
ranks must equal the undecomposed run BIT FOR BIT (reproducibility across GPU counts)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, make_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, mode, tmp_path, nsteps, scheme, adaptive):
    out = str(tmp_path / f"res_{mode}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_dist_worker.py"), mode, out, str(nsteps), str(scheme), "1" if adaptive else "0"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return [np.load(f"{out}.{k}.npz") for k in range(world)]


def _reference(mode, nsteps, scheme, adaptive):
    from swe_fvm_b200 import TriangMesh
    from oracle.oracle import Oracle
    if mode == "strips":
        mesh, case, v0 = make_case("classic_thacker", 24, quad_n=4)
    else:
        bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
        mesh, case, v0 = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    o = Oracle(mesh)
    o.set_state(v0)
    o.run(scheme, 1, 2, nsteps, 0.0 if adaptive else 2e-3, 1e-3)
    return o.get_state(), o.cfl_dt()


@pytest.mark.parametrize("world,mode,scheme,adaptive", [
    (2, "strips", 1, True), (3, "strips", 2, False), (2, "general", 1, True), (3, "general", 0, True)])
def test_decomposed_run_is_bitwise_identical(tmp_path, world, mode, scheme, adaptive):
    nsteps = 25
    res = _run(world, mode, tmp_path, nsteps, scheme, adaptive)
    want, want_dt = _reference(mode, nsteps, scheme, adaptive)
    seen = np.zeros(len(want), dtype=int)
    for r in res:
        np.testing.assert_array_equal(r["state"], want[r["gids"]])
        seen[r["gids"]] += 1
        if adaptive:
            assert float(r["dt"]) == want_dt  # global min all-reduce => identical dt on every rank
        assert int(r["exchanges"]) == nsteps * (scheme + 1)
        assert int(r["nsend"]) > 0 and int(r["nrecv"]) > 0
    assert (seen == 1).all()  # every cell owned exactly once


def test_strip_plans_are_consistent():
    """Rank r's receive list from r+1 addresses the same global cells, in the same order, as rank r+1's
    send list to r; every cell is owned exactly once; the CFL edge mask covers exactly the edges touching
    owned cells; the ordering classes are what swe_dist's launch order relies on."""
    from swe_fvm_b200 import dist as swd
    n, world = 16, 4
    plans = [swd.Plan.struct(r, world, n, n, 4.0 / n) for r in range(world)]
    assert sum(p.n_owned for p in plans) == 4 * n * n
    seen = np.zeros(4 * n * n, int)
    for p in plans:
        seen[p.global_cells[p.owned]] += 1
    assert (seen == 1).all()
    for r in range(world - 1):
        up = [p for p in plans[r].peers if p[0] == r + 1][0]
        dn = [p for p in plans[r + 1].peers if p[0] == r][0]
        np.testing.assert_array_equal(plans[r].global_cells[up[1]], plans[r + 1].global_cells[dn[2]])
        np.testing.assert_array_equal(plans[r].global_cells[up[2]], plans[r + 1].global_cells[dn[1]])
    _check_plan_invariants(plans[1])


def _check_plan_invariants(p):
    et, tt = p.mesh.edge_elements, p.mesh.element_neighbours
    touch = p.owned[et[:, 0]] | np.where(et[:, 1] >= 0, p.owned[np.maximum(et[:, 1], 0)], False)
    np.testing.assert_array_equal(p.cfl_mask.astype(bool), touch)
    halo = np.zeros(p.mesh.nt, bool)
    halo[p.recv_list()] = True
    assert (~p.owned == halo).all()                       # every non-owned cell is received from some peer
    sent = np.zeros(p.mesh.nt, bool)
    sent[p.send_list()] = True
    assert not (sent & ~p.owned).any()                    # only owned cells are sent
    dep = np.array([any(j >= 0 and halo[j] for j in tt[i]) for i in range(p.mesh.nt)])
    cls = p.classes
    np.testing.assert_array_equal(cls == 3, halo)         # class 3: halo cells
    np.testing.assert_array_equal(cls == 2, dep & ~halo)  # class 2: owned, stencil touches a halo cell
    np.testing.assert_array_equal(cls == 1, sent & ~dep)  # class 1: sent, stencil free of halo cells
    assert not ((cls == 2) & ~sent).any()                 # halo-dependent owned cells are always sent


def test_general_plans_derive_send_lists_without_communication():
    """Any mesh + partition vector: each rank derives its send lists by repeating the peers' ring growth, and
    they pair up exactly with the peers' receive lists."""
    from swe_fvm_b200 import TriangMesh
    from swe_fvm_b200 import dist as swd
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    world = 3
    part = bowl.partition_rcb(world)
    plans = [swd.Plan.from_mesh(r, world, bowl, part) for r in range(world)]
    seen = np.zeros(bowl.nt, int)
    for p in plans:
        seen[p.global_cells[p.owned]] += 1
        _check_plan_invariants(p)
        assert 0 < (p.classes >= 2).sum() < 0.3 * p.mesh.nt
    assert (seen == 1).all()
    for r in range(world):
        for peer, s, rcv in plans[r].peers:
            back = [q for q in plans[peer].peers if q[0] == r][0]
            np.testing.assert_array_equal(plans[r].global_cells[s], plans[peer].global_cells[back[2]])
            np.testing.assert_array_equal(plans[r].global_cells[rcv], plans[peer].global_cells[back[1]])
