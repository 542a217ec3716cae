#!/usr/bin/env python
"""Throughput of BASELINE.json's configs[0..3] on one GPU next to the CPU oracle (1 thread):
fills the results table of BASELINE.md §6. Run on the GPU box: python scripts/bench_configs.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from swe_fvm_b200 import Case, StructTriangMesh, TriangMesh  # noqa: E402
from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def gpu_rate(mesh, v0, scheme, steps, reorder, dt):
    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, reorder=reorder)
    td = TimeDisc(sd)
    Solvers.run(td, scheme, 3, dt=1e-6)
    sd.synchronize()
    dt0 = td.CFLdt() if dt is None else dt
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    Solvers.run(td, scheme, steps, dt=0.0 if dt is None else dt, dt0=dt0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    d = sd.diagnostics()
    return mesh.nt * steps / (ms * 1e-3), ms / steps, d["wet_cells"] / mesh.nt


def cpu_rate(mesh, v0, scheme, steps, dt):
    o = Oracle(mesh)
    o.set_state(v0)
    sc = {"euler": 0, "ssprk2": 1, "ssprk3": 2}[scheme]
    o.step(sc, 1, 2, 1e-6)
    d = o.cfl_dt() if dt is None else dt
    t = time.perf_counter()
    o.run(sc, 1, 2, steps, 0.0 if dt is None else dt, d)
    return mesh.nt * steps / (time.perf_counter() - t)


def main():
    def lake():
        m = StructTriangMesh(71, 71, 4 / 71); c = Case("lake_at_rest", 2, 2, 4); c.set_bathymetry(m)
        return m, c.initial_state(m)

    def thacker(n, q):
        m = StructTriangMesh(n, n, 4 / n); c = Case("classic_thacker", 2, 2, 4); c.set_bathymetry(m)
        return m, c.initial_state(m, q)

    def wet(n):
        m = StructTriangMesh(n, n, 4 / n); c = Case("fully_wet", 2, 2, 4); c.set_bathymetry(m)
        return m, c.initial_state(m, 1)

    def bowl(levels):
        m = TriangMesh.from_gmsh(os.path.join(ROOT, "tests", "golden", "bowl.msh"))
        for _ in range(levels):
            m = m.refine()
        c = Case("bowl_hump", 4, 4, 8, level=3.0, amp=0.5); c.set_bathymetry(m)
        return m, c.initial_state(m, 2)

    rows = [
        ("0 LakeAtRest 20k, Euler dt=1e-3", lake, "euler", 2000, False, 1e-3, 50),
        ("1 ClassicThacker 1M", lambda: thacker(512, 2), "ssprk2", 200, False, None, 5),
        ("2 Bowl bowl.msh x4 (3.8M), caller order", lambda: bowl(4), "ssprk2", 100, False, None, 3),
        ("2 Bowl bowl.msh x4 (3.8M), Hilbert order", lambda: bowl(4), "ssprk2", 100, True, None, 0),
        ("- bowl.msh unrefined (15k), caller order", lambda: bowl(0), "ssprk2", 2000, False, None, 0),
        ("- bowl.msh unrefined (15k), Hilbert order", lambda: bowl(0), "ssprk2", 2000, True, None, 0),
        ("3 Thacker 64M (mostly dry)", lambda: thacker(4096, 1), "ssprk2", 20, False, None, 0),
        ("3 fully wet 64M", lambda: wet(4096), "ssprk2", 20, False, None, 0),
    ]
    for name, make, scheme, steps, reorder, dt, cpu_steps in rows:
        mesh, v0 = make()
        g, ms, wetf = gpu_rate(mesh, v0, scheme, steps, reorder, dt)
        cpu = cpu_rate(mesh, v0, scheme, cpu_steps, dt) if cpu_steps else None
        print(json.dumps(dict(config=name, cells=mesh.nt, scheme=scheme, wet_fraction=wetf, gpu_cell_updates_per_s=g,
                              ms_per_step=ms, cpu_oracle_1thread_cell_updates_per_s=cpu)), flush=True)
        del mesh, v0


if __name__ == "__main__":
    main()
