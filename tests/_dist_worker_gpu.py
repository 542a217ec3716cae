"""Worker of tests/test_gpu_dist.py::test_two_gpus_*: one rank (= one GPU, one process) of a torchrun job
running swe_dist (peer memory over CUDA IPC); torch.distributed is only the bootstrap all-gather."""
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

    if mode == "strips":
        n = 64
        plan = swd.Plan.struct(rank, world, n, n, 4.0 / n)
        case = Case("classic_thacker", 2.0, 2.0, 4.0)
    else:
        g = TriangMesh.from_gmsh(os.path.join(ROOT, "tests", "golden", "bowl.msh"))
        case = Case("bowl_hump", 4.0, 4.0, 8.0, level=3.0, amp=0.5)
        case.set_bathymetry(g)
        plan = swd.Plan.from_mesh(rank, world, g, g.partition_rcb(world))
    case.set_bathymetry(plan.mesh)
    v0 = case.initial_state(plan.mesh, quad_n=4)
    ds = swd.DistSolver(plan, device=local_rank, reorder=(mode != "strips"), wait_timeout_s=30.0)
    ds.sd.SetVolField(v0)
    ds.exchange()
    ds.run(scheme, nsteps, dt=0.0 if adaptive else 2e-3, dt0=1e-3)
    ds.synchronize()
    gids, st = ds.owned_state()
    np.savez(f"{out}.{rank}.npz", gids=gids, state=st, hash=np.uint64(ds.state_hash()), dt=ds.cfl_dt())
    dist.barrier()
    ds.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
