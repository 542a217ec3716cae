"""Multi-GPU path on the device, through the swe_dist_* C-ABI (csrc/swe_dist.cuh).

(1) swe_dist_group_create with every rank on ONE GPU: the complete transport runs for real — ordering
    classes, class-split K1 / K4 launches, k_halo_pack_signal (stores + flag), k_halo_wait_unpack (spin on the
    flags), the peer-memory CFL minimum (k_min_push / k_min_pull) — only the CUDA-IPC mapping is replaced by
    direct pointers. Owned cells must equal the single-context run BIT FOR BIT, the order-independent
    state hash and the final dt must agree.
(2) With >= 2 GPUs: one process per GPU under torchrun (NCCL only as the bootstrap all-gather), CUDA IPC."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, make_case

pytestmark = pytest.mark.gpu


def _single(mesh, v0, scheme, nsteps, dt=0.0, cor=0.0):
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=cor)
    td = TimeDisc(sd)
    Solvers.run(td, scheme, nsteps, dt=dt, dt0=1e-3)
    sd.synchronize()
    return sd.GetVolField(), sd.state_hash(), td.CFLdt()


def _group_run(plans, case, scheme, nsteps, reorder, overlap, dt=0.0, cor=0.0):
    from swe_fvm_b200 import dist as swd
    v0s = []
    for p in plans:
        case.set_bathymetry(p.mesh)
        v0s.append(case.initial_state(p.mesh, quad_n=4))
    grp = swd.DistGroup(plans, [0] * len(plans), cor=cor, reorder=reorder, overlap=overlap, wait_timeout_s=20.0)
    for sd, v0 in zip(grp.sds, v0s):
        sd.SetVolField(v0)
    grp.exchange()
    grp.run(scheme, nsteps, dt=dt, dt0=1e-3)
    grp.synchronize()
    return grp


@pytest.mark.parametrize("world,scheme,overlap,reorder", [(2, "ssprk2", True, True), (4, "ssprk3", True, False),
                                                          (3, "euler", False, True), (2, "ssprk2", True, False)])
def test_group_strips_on_one_gpu_bitwise(world, scheme, overlap, reorder):
    from swe_fvm_b200 import dist as swd
    n = 48
    mesh, case, v0 = make_case("classic_thacker", n, quad_n=4)
    want, want_hash, want_dt = _single(mesh, v0, scheme, 30, cor=0.2)
    plans = [swd.Plan.struct(r, world, n, n, 4.0 / n) for r in range(world)]
    grp = _group_run(plans, case, scheme, 30, reorder, overlap, cor=0.2)
    seen = np.zeros(mesh.nt, int)
    for g, st in grp.owned_states():
        np.testing.assert_array_equal(st, want[g])
        seen[g] += 1
    assert (seen == 1).all()
    assert grp.state_hash() == want_hash
    for r in range(world):
        assert grp.cfl_dt(r) == want_dt   # the peer-memory minimum is exact: same dt on every rank as on one GPU
    grp.close()


def test_group_rcb_partitions_on_one_gpu_bitwise():
    from swe_fvm_b200 import TriangMesh
    from swe_fvm_b200 import dist as swd
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    mesh, case, v0 = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    want, want_hash, want_dt = _single(mesh, v0, "ssprk2", 30)
    world = 3
    part = bowl.partition_rcb(world)
    plans = [swd.Plan.from_mesh(r, world, bowl, part) for r in range(world)]
    grp = _group_run(plans, case, "ssprk2", 30, True, True)
    seen = np.zeros(mesh.nt, int)
    for g, st in grp.owned_states():
        np.testing.assert_array_equal(st, want[g])
        seen[g] += 1
    assert (seen == 1).all()
    assert grp.state_hash() == want_hash and grp.cfl_dt(0) == want_dt
    grp.close()


def test_group_fixed_dt_steps_and_hash_detects_a_flipped_bit():
    from swe_fvm_b200 import dist as swd
    n = 32
    mesh, case, v0 = make_case("classic_thacker", n, quad_n=4)
    want, want_hash, _ = _single(mesh, v0, "ssprk2", 10, dt=2e-3)
    plans = [swd.Plan.struct(r, 2, n, n, 4.0 / n) for r in range(2)]
    grp = _group_run(plans, case, "ssprk2", 10, True, True, dt=2e-3)
    assert grp.state_hash() == want_hash
    st = grp.sds[0].GetVolField()
    i = int(np.nonzero(plans[0].owned)[0][5])
    st[i, 0] = np.nextafter(st[i, 0], np.inf)
    grp.sds[0].SetVolField(st)
    assert grp.state_hash() != want_hash
    grp.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["strips", "general"])
def test_two_gpus_bitwise(tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from swe_fvm_b200 import TriangMesh
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker_gpu.py"), mode, out, "30", "1", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    if mode == "strips":
        mesh, case, v0 = make_case("classic_thacker", 64, quad_n=4)
    else:
        bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
        mesh, case, v0 = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    want, want_hash, want_dt = _single(mesh, v0, "ssprk2", 30)
    seen = np.zeros(mesh.nt, int)
    tot = 0
    for k in range(2):
        res = np.load(f"{out}.{k}.npz")
        np.testing.assert_array_equal(res["state"], want[res["gids"]])
        seen[res["gids"]] += 1
        tot = (tot + int(res["hash"])) % (1 << 64)
        assert float(res["dt"]) == want_dt
    assert (seen == 1).all() and tot == want_hash
