#!/usr/bin/env python
"""configs[1] (ClassicThacker on StructTriangMesh(512): 1M cells, SSPRK2, dt = CFLdt) in isolation: ms/step with and
without CUDA-graph replay. Under `ncu --metrics gpu__time_duration.sum` (use --graph 0) the launch list shows how much
of a step is kernel time and how much is launch gaps.  python scripts/probe_small.py [--n 512] [--steps 200] [--graph -1|0|1]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from swe_fvm_b200 import Case, StructTriangMesh  # noqa: E402
from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--graph", type=int, default=-1)
    ap.add_argument("--reorder", type=int, default=1)
    a = ap.parse_args()
    m = StructTriangMesh(a.n, a.n, 4 / a.n)
    c = Case("classic_thacker", 2, 2, 4)
    c.set_bathymetry(m)
    sd = SpaceDisc("hllc", "einfeldt", m, c.initial_state(m, 2), reorder=bool(a.reorder))
    sd.set_option("graph", a.graph)
    td = TimeDisc(sd)
    Solvers.run(td, "ssprk2", 5, dt=1e-6)
    sd.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sd.launch_count()
    torch.cuda.synchronize()
    e0.record()
    Solvers.run(td, "ssprk2", a.steps, dt=0.0, dt0=td.CFLdt())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"cells": m.nt, "graph": a.graph, "reorder": a.reorder, "ms_per_step": ms, "cell_updates_per_s": m.nt / (ms * 1e-3),
                      "launches_per_step": (sd.launch_count() - l0) / a.steps}))


if __name__ == "__main__":
    main()
