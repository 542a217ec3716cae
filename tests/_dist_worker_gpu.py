"""Worker of tests/test_gpu_dist.py::test_nccl_*: one rank (= one GPU) of an NCCL job running the
product's DistributedSolver on the device; writes the owned cells' final state per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, out, nsteps, scheme, adaptive = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    from swe_fvm_b200 import Case, TriangMesh
    from swe_fvm_b200 import dist as swd
    from swe_fvm_b200.solver import SpaceDisc

    if mode == "strips":
        n = 64
        dec = swd.decompose_strips(n, n, 4.0 / n, rank, world)
        case = Case("classic_thacker", 2.0, 2.0, 4.0)
        lo = max(swd.strip_rows(n, world)[rank][0] - swd.HALO_ROWS, 0)
        gids = np.arange(dec.mesh.nt) + lo * 4 * n
    else:
        g = TriangMesh.from_gmsh(os.path.join(ROOT, "tests", "golden", "bowl.msh"))
        case = Case("bowl_hump", 4.0, 4.0, 8.0, level=3.0, amp=0.5)
        case.set_bathymetry(g)
        part = g.partition_rcb(world)
        dec = swd.decompose_general(g, part, rank, world)
        gids = dec.global_cells
    case.set_bathymetry(dec.mesh)
    v0 = case.initial_state(dec.mesh, quad_n=4)
    sd = SpaceDisc("hllc", "einfeldt", dec.mesh, v0, device=local_rank, reorder=(mode != "strips"),
                   cell_class=dec.cell_classes())
    sd.set_stream(torch.cuda.current_stream().cuda_stream)
    local = swd.GpuLocal(sd, has_classes=True)
    transport = os.environ.get("SWE_HALO", "nccl")
    solver = swd.DistributedSolver(dec, local, transport=transport)
    assert solver.halo.transport == transport
    solver.run(scheme, nsteps, None if adaptive else 2e-3, dt0=1e-3)
    sd.synchronize()
    assert not local.p2p_error()
    st = sd.GetVolField()
    np.savez(f"{out}.{rank}.npz", gids=gids[dec.owned], state=st[dec.owned], minlen=float(local.min_len_tensor().item()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
