"""Worker of tests/test_dist_cpu.py: one rank of a world_size-N gloo job. Runs the product's
C++ decomposition plan + the test-side stage loop with the CPU oracle as the local solver and writes the owned
cells' final state (with global ids) to <out>.<rank>.npz."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode, out, nsteps, scheme, adaptive = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from swe_fvm_b200 import Case, StructTriangMesh, TriangMesh
    from swe_fvm_b200 import dist as swd
    from oracle.oracle import Oracle
    from dist_helpers import OracleLocal

    if mode == "strips":
        n = 24
        plan = swd.Plan.struct(rank, world, n, n, 4.0 / n)
        case = Case("classic_thacker", 2.0, 2.0, 4.0)
    else:
        g = TriangMesh.from_gmsh(os.path.join(ROOT, "tests", "golden", "bowl.msh"))
        case = Case("bowl_hump", 4.0, 4.0, 8.0, level=3.0, amp=0.5)
        case.set_bathymetry(g)
        plan = swd.Plan.from_mesh(rank, world, g, g.partition_rcb(world))
    gids = plan.global_cells
    case.set_bathymetry(plan.mesh)
    v0 = case.initial_state(plan.mesh, quad_n=4)
    o = Oracle(plan.mesh)
    o.set_state(v0)
    local = OracleLocal(o)
    from dist_helpers import EmulatedSolver
    solver = EmulatedSolver(plan, local)
    solver.run(scheme, nsteps, None if adaptive else 2e-3, dt0=1e-3)
    st = o.get_state()
    np.savez(f"{out}.{rank}.npz", gids=gids[plan.owned], state=st[plan.owned], dt=local.dt,
             minlen=float(local._ml[0]), exchanges=solver.exchanges, nsend=solver.nsend, nrecv=solver.nrecv)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
