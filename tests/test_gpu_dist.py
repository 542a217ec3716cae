"""Multi-GPU path on the device. (1) P partitions emulated on ONE GPU: P contexts stepped in
lockstep with the product's pack/unpack kernels, CFL edge mask and device-resident dt — owned
cells must equal the single-context run bit for bit. (2) With >= 2 GPUs: the real NCCL job."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, make_case

pytestmark = pytest.mark.gpu

STAGES = {0: [(0.0, 1.0, 1.0)], 1: [(0.0, 1.0, 1.0), (0.5, 0.5, 0.5)],
          2: [(0.0, 1.0, 1.0), (0.75, 0.25, 0.25), (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0)]}


def _emulate(decs, gids, case, scheme, nsteps, reorder):
    """Lockstep emulation of DistributedSolver.step for all ranks inside one process."""
    import torch
    from swe_fvm_b200 import dist as swd
    from swe_fvm_b200.solver import SpaceDisc
    locs, sds, bufs = [], [], []
    for d in decs:
        case.set_bathymetry(d.mesh)
        v0 = case.initial_state(d.mesh, quad_n=4)
        sd = SpaceDisc("hllc", "einfeldt", d.mesh, v0, reorder=reorder)
        L = swd.GpuLocal(sd)
        L.set_cfl_edge_mask(d.cfl_edge_mask())
        L.set_halo_lists(d.send_list(), d.recv_list())
        ns, nr = len(d.send_list()), len(d.recv_list())
        bufs.append((L.alloc(3 * max(ns, 1)), L.alloc(3 * max(nr, 1))))
        L.set_dt(1e-3)
        locs.append(L)
        sds.append(sd)

    def exchange():
        for L, (sb, rb) in zip(locs, bufs):
            L.pack(sb)
        torch.cuda.synchronize()
        for r, d in enumerate(decs):
            ro = 0
            for peer, s, rcv in d.peers:
                # find my segment in the peer's send buffer
                so = 0
                for p2, s2, r2 in decs[peer].peers:
                    if p2 == r:
                        bufs[r][1][3 * ro:3 * (ro + len(rcv))] = bufs[peer][0][3 * so:3 * (so + len(s2))]
                        assert len(s2) == len(rcv)
                        np.testing.assert_array_equal(gids[peer][s2], gids[r][rcv])
                        break
                    so += len(s2)
                ro += len(rcv)
        for L, (sb, rb) in zip(locs, bufs):
            L.unpack(rb)

    for _ in range(nsteps):
        st = STAGES[scheme]
        for k, (a0, a1, coef) in enumerate(st):
            for L in locs:
                L.compute_interface_values()
                L.compute_fluxes()
            if k == len(st) - 1:  # global min all-reduce
                mn = min(float(L.min_len_tensor().item()) for L in locs)
                for L in locs:
                    L.min_len_tensor().fill_(mn)
            for L in locs:
                if k == 0 and len(st) > 1:
                    L.save_state()
                L.stage_update(a0, a1, coef, None)
            exchange()
        for L in locs:
            L.advance_dt(None)
    out = []
    for sd, d, g in zip(sds, decs, gids):
        sd.synchronize()
        out.append((g[d.owned], sd.GetVolField()[d.owned]))
    return out


def _single(mesh, v0, scheme, nsteps):
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0)
    Solvers.run(TimeDisc(sd), scheme, nsteps, dt=0.0, dt0=1e-3)
    return sd.GetVolField()


@pytest.mark.parametrize("world,scheme", [(2, 1), (4, 2)])
def test_emulated_strips_bitwise(world, scheme):
    from swe_fvm_b200 import dist as swd
    n = 48
    mesh, case, v0 = make_case("classic_thacker", n, quad_n=4)
    want = _single(mesh, v0, scheme, 30)
    decs = [swd.decompose_strips(n, n, 4.0 / n, r, world) for r in range(world)]
    gids = [np.arange(d.mesh.nt) + max(swd.strip_rows(n, world)[r][0] - swd.HALO_ROWS, 0) * 4 * n for r, d in enumerate(decs)]
    seen = np.zeros(mesh.nt, int)
    for g, st in _emulate(decs, gids, case, scheme, 30, reorder=False):
        np.testing.assert_array_equal(st, want[g])
        seen[g] += 1
    assert (seen == 1).all()


def test_emulated_rcb_partitions_bitwise_with_reordering():
    from swe_fvm_b200 import TriangMesh
    from swe_fvm_b200 import dist as swd
    bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
    mesh, case, v0 = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    want = _single(mesh, v0, 1, 30)
    world = 3
    part = bowl.partition_rcb(world)
    wants = [None] * world
    decs = []
    # emulate all_gather_object: first pass collects every rank's wish list
    for r in range(world):
        sub = bowl.extract(part, r, swd.HALO_LAYERS)
        gc, owner = np.array(sub.global_cells), np.array(sub.cell_owner)
        wants[r] = {int(q): gc[owner == q] for q in np.unique(owner) if q != r}
    for r in range(world):
        decs.append(swd.decompose_general(bowl, part, r, world, all_gather_object=lambda w: wants))
    gids = [np.array(d.global_cells) for d in decs]
    seen = np.zeros(mesh.nt, int)
    for g, st in _emulate(decs, gids, case, 1, 30, reorder=True):
        np.testing.assert_array_equal(st, want[g])
        seen[g] += 1
    assert (seen == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("mode", ["strips", "general"])
def test_nccl_two_gpus_bitwise(tmp_path, mode, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from swe_fvm_b200 import TriangMesh
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker_gpu.py"), mode, out, "30", "1", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SWE_HALO=transport))
    assert r.returncode == 0, r.stderr[-3000:]
    if mode == "strips":
        mesh, case, v0 = make_case("classic_thacker", 64, quad_n=4)
    else:
        bowl = TriangMesh.from_gmsh(os.path.join(GOLDEN, "bowl.msh"))
        mesh, case, v0 = make_case("bowl_hump", mesh=bowl, level=3.0, amp=0.5)
    want = _single(mesh, v0, 1, 30)
    seen = np.zeros(mesh.nt, int)
    for k in range(2):
        res = np.load(f"{out}.{k}.npz")
        np.testing.assert_array_equal(res["state"], want[res["gids"]])
        seen[res["gids"]] += 1
    assert (seen == 1).all()
