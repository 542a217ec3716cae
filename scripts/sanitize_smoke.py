"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the
step incl. the part-wet pass, both orderings, taps, halo pack/unpack, diagnostics, device cases."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from swe_fvm_b200 import Case, StructTriangMesh, TriangMesh  # noqa: E402
from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc  # noqa: E402

for reorder in (False, True):
    mesh = StructTriangMesh(24, 24, 4 / 24)
    case = Case("classic_thacker", 2, 2, 4)
    case.set_bathymetry(mesh)
    v0 = case.initial_state(mesh, 4)
    cls = (np.arange(mesh.nt) % 7 == 0).astype(np.uint8)
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0.2, reorder=reorder, taps=True, cell_class=cls)
    td = TimeDisc(sd)
    for scheme in ("euler", "ssprk2", "ssprk3"):
        Solvers.run(td, scheme, 5, dt=2e-3)
    Solvers.run(td, "ssprk2", 5, dt=0.0, dt0=1e-3)
    sd.ComputeInterfaceValues(); sd.ComputeFluxes()
    sd.GetEdgField(); sd.GetSrcField(); sd.GetFluxes(); sd.node_max_w(); sd.draining_dt(); sd.cell_class()
    sd._call("swe_compute_interface_values_class", 0, 1, 0)
    sd._call("swe_compute_interface_values_class", 1, 0, 1)
    print(sd.diagnostics(), sd.case_l2_error(case, sd.time()))
    sd.set_case_state(case, 3)
# round 2: dry-region instantiations (forced on, graph replay and plain launches), the three forms of the reconstruction
# kernel, fused draining dt, CFL-free flux kernel, checkpoint, host-buffer pipeline, single-process multi-GPU group
mesh = StructTriangMesh(40, 40, 4 / 40)
case = Case("classic_thacker", 2, 2, 4)
case.set_bathymetry(mesh)
v0 = case.initial_state(mesh, 4)
for opts in (dict(dry_skip=1, graph=0), dict(dry_skip=1, graph=1), dict(k1_tiled=1), dict(k1_tiled=2), dict(fused_drain=1),
             dict(skip_cfl=0, dry_skip=0)):
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, reorder=True)
    for k, v in opts.items():
        sd.set_option(k, v)
    td = TimeDisc(sd)
    for scheme in ("euler", "ssprk2", "ssprk3"):
        Solvers.run(td, scheme, 6, dt=2e-3)
        getattr(Solvers, {"euler": "Euler", "ssprk2": "SSPRK2", "ssprk3": "SSPRK3"}[scheme])(td, 2e-3)
    sd.rhs(1e-3)
    sd.synchronize()
    print(opts, sd.diagnostics()["mass"])
from swe_fvm_b200 import dist as swd  # noqa: E402
plans = [swd.Plan.struct(r, 2, 16, 32, 4 / 16) for r in range(2)]
gcase = Case("classic_thacker", 2, 4, 4)
v0s = []
for p in plans:
    gcase.set_bathymetry(p.mesh)
    v0s.append(gcase.initial_state(p.mesh, quad_n=2))
grp = swd.DistGroup(plans, [0, 0], reorder=True, overlap=True, wait_timeout_s=60.0)
for gsd, gv in zip(grp.sds, v0s):
    gsd.SetVolField(gv)
grp.exchange()
grp.run("ssprk2", 6, dt=0.0, dt0=1e-3)
grp.synchronize()
print("group ok", grp.state_hash())
grp.close()
bowl = TriangMesh.from_gmsh(os.path.join(ROOT, "tests", "golden", "bowl.msh"))
case = Case("bowl_hump", 4, 4, 8, level=3.0, amp=0.5)
case.set_bathymetry(bowl)
sd = SpaceDisc("hll", "rusanov", bowl, case.initial_state(bowl), reorder=True)
Solvers.run(TimeDisc(sd), "ssprk2", 10, dt=1e-3)
print("sanitize run ok", sd.diagnostics()["mass"])
