#!/usr/bin/env python
"""bench.py — fp64 cell-updates/s of the SWE_FVM explicit time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n NI] [--case fully_wet|thacker]

A "step" is one SSPRK2 HLLC<Einfeldt> time step (2 stages) of every cell of the workload:
  N = 1 : configs[3] — synthetic StructTriangMesh(4096, 4096, 4/4096), 67 108 864 cells, fully-wet
          variant (the wet-cell fraction is printed), dt = CFLdt() of the previous step.
  N > 1 : weak scaling — rank r owns a 4096 x 4096 strip of StructTriangMesh(4096, 4096 N, h) plus
          3 halo rows per open side; one NCCL halo exchange per stage, one min all-reduce per step.
`value` = cells * K / (max over ranks of the CUDA-event time of K steps), state resident in HBM.
`e2e`   = same, through the host-buffer C-ABI calls: every step uploads the state from pinned
          host memory (swe_set_state_async), steps, and downloads it again (swe_get_state_async).
`--impl reference` times the CPU oracle (the upstream code does not build here)
with all host threads on a bounded sample of the same workload. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fp64 cell-updates/s"
UNIT = "cell-updates/s"
B_PER_CELL_UPDATE_SSPRK2 = 1488.0  # 744 B per cell-stage x 2 stages
# algorithmic bytes per cell per launch of each kernel in THIS implementation's decomposition
# (every array element counted once per kernel that needs it, structured
# ratios Ne/Nt = 1.5, Nn/Nt = 0.5)
KERNEL_BYTES_PER_CELL = {
    "k_reconstruct": 185.0,  # W 24 + tt,tp 24 + cgeo 32 + nodes 16 | ceh,ceu,cev 72 + cgx,cgy 16 + cls 1
    "k_partwet2": 0.0,       # work list only (O(sqrt N) part-wet cells)
    "k_flux": 156.0,         # per edge: slots 8 + n 16 + dmin 8 + two sides 48 | F 24; x 1.5 edges per cell
    "k_drain": 56.0,         # te 12 + F0 12 + w 8 + cb 8 + area 8 | dti 8
    "k_update": 220.0,       # W 24 + (U0 24 on the 2nd stage) + te,tt 24 + F 36 + dti 8 + ceh 24 + grad 16 + n 24 +
                             # L 12 + area 8 + cb 8 | W 24  -> 208 (stage 1) / 232 (stage 2), mean 220
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# profiles/r1_v3_ncu_full_summary.csv (64M-cell fully-wet workload); None for other workloads
NCU_TRAFFIC_BYTES_64M = {"k_reconstruct": 12.752e9, "k_flux": 10.571e9, "k_drain": 3.751e9, "k_update": 13.980e9}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def build_case(kind: str, mesh, mid_y: float, length: float):
    from swe_fvm_b200 import Case
    if kind == "thacker":
        case = Case("classic_thacker", 2.0, mid_y, length)
    else:
        case = Case("fully_wet", 2.0, mid_y, length)
    case.set_bathymetry(mesh)
    return case, case.initial_state(mesh, quad_n=1)


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample (cpu_baseline of the GPU line, and --impl reference)
# ----------------------------------------------------------------------------------------------
def time_oracle(n: int, case_kind: str, steps: int, warmup: int, threads: int):
    from swe_fvm_b200 import StructTriangMesh
    from oracle.oracle import Oracle
    mesh = StructTriangMesh(n, n, 4.0 / n)
    case, v0 = build_case(case_kind, mesh, 2.0, 4.0)
    o = Oracle(mesh, threads=threads)
    o.set_state(v0)
    o.step(1, 1, 2, 1e-4)  # primes min_len
    dt = o.cfl_dt()
    for _ in range(max(warmup - 1, 0)):
        o.step(1, 1, 2, dt)
        dt = o.cfl_dt()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(1, 1, 2, dt)
        dt = o.cfl_dt()
    el = time.perf_counter() - t0
    return mesh.nt * steps / el, el, mesh.nt, o.threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build()
    threads = os.cpu_count() or 1
    # size the sample so that (steps + warmup) steps finish within ~2 minutes
    rate1, _, _, _ = time_oracle(128, args.case, 2, 1, threads)
    budget_cells = rate1 * 90.0 / max(args.steps + args.warmup, 1)
    n = 128
    for cand in (2048, 1024, 512, 256):
        if 4 * cand * cand <= budget_cells:
            n = cand
            break
    v, el, cells, thr = time_oracle(n, args.case, args.steps, args.warmup, threads)
    sample = (f"StructTriangMesh({n},{n},4/{n}) = {cells} cells of the {args.case} workload, {args.steps} SSPRK2 "
              f"HLLC<Einfeldt> steps, dt=CFLdt, OpenMP oracle (Jacobi semantics)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, max(args.gpus, 1)), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "upstream SWE_FVM does not build (Eigen fetched at configure time, HEAD mid-refactor); the CPU "
                "oracle restating it is timed instead",
    }
    print(json.dumps(line), flush=True)


def workload_name(args, world):
    n = args.n
    if args.global_n:
        g = args.global_n
        return (f"configs[4]: synthetic StructTriangMesh({g},{g},4/{g}) = {4 * g * g} cells split into {world} strips "
                f"(strong scaling), {args.case} variant, HLLC<Einfeldt>, SSPRK2, dt=CFLdt")
    if world == 1:
        return (f"configs[3]: synthetic StructTriangMesh({n},{n},4/{n}) = {4 * n * n} cells, {args.case} variant, "
                f"HLLC<Einfeldt>, SSPRK2, dt=CFLdt")
    return (f"configs[4]-style weak scaling: StructTriangMesh({n},{n * world},4/{n}) = {4 * n * n * world} cells in "
            f"{world} strips of {4 * n * n} cells, {args.case} variant, HLLC<Einfeldt>, SSPRK2, dt=CFLdt")


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        import __graft_entry__ as g
        g.build()
    if world > 1:
        dist.barrier()
    from swe_fvm_b200 import StructTriangMesh
    from swe_fvm_b200 import dist as swd
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc

    n, h = args.n, 4.0 / args.n
    t_setup = time.perf_counter()
    if args.global_n:  # strong scaling: one fixed global mesh, split into `world` strips
        n, h = args.global_n, 4.0 / args.global_n
        if world == 1:
            mesh, dec, length_y = StructTriangMesh(n, n, h), None, 4.0
            n_owned = mesh.nt
        else:
            dec = swd.decompose_strips(n, n, h, rank, world)
            mesh, n_owned, length_y = dec.mesh, dec.n_owned, 4.0
    elif world == 1:
        mesh = StructTriangMesh(n, n, h)
        dec = None
        n_owned = mesh.nt
        length_y = 4.0
    else:
        dec = swd.decompose_strips(n, n * world, h, rank, world)
        mesh = dec.mesh
        n_owned = dec.n_owned
        length_y = 4.0 * world
    case, v0 = build_case(args.case, mesh, 0.5 * length_y, 4.0)
    sd = SpaceDisc("hllc", "einfeldt", mesh, None, device=local_rank, reorder=args.reorder,
                   cell_class=(dec.cell_classes() if dec is not None else None))
    sd.set_stream(torch.cuda.current_stream().cuda_stream)
    for kv in args.opt:
        k, v = kv.split("=")
        sd.set_option(k, int(v))
    td = TimeDisc(sd)
    pin_in = torch.from_numpy(v0).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    sd.set_state_async(pin_in.data_ptr())
    sd.synchronize()
    d0 = sd.diagnostics()
    wet_frac = d0["wet_cells"] / mesh.nt
    t_setup = time.perf_counter() - t_setup

    SSPRK2 = 1
    if dec is not None:
        local = swd.GpuLocal(sd, has_classes=True)
        solver = swd.DistributedSolver(dec, local, overlap=not args.no_overlap, transport=args.halo)

        def run_steps(k):
            solver.run(SSPRK2, k, None, dt0=dt_first[0])
    else:
        def run_steps(k):
            Solvers.run(td, "ssprk2", k, dt=0.0, dt0=dt_first[0])

    # prime the CFL dt with one tiny step, then warm up
    dt_first = [1e-6]
    run_steps(1)
    sd.synchronize()
    if dec is not None:
        mn = local.min_len_tensor().clone()
        dt_first[0] = 0.15 * float(mn.item())
    else:
        dt_first[0] = td.CFLdt()
    run_steps(max(args.warmup, 1))
    sd.synchronize()
    dt_first[0] = 0.15 * float(local.min_len_tensor().item()) if dec is not None else td.CFLdt()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, state resident in HBM ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sd.kernel_timing(True)
    launches0 = sd.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    launches = sd.launch_count() - launches0
    ktimes = sd.kernel_times()
    sd.kernel_timing(False)
    sd.synchronize()  # raises if a non-finite state was produced
    d1 = sd.diagnostics()

    # ---- e2e: host buffers in and out every step (C-ABI with pinned host memory) ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    dt_e2e = dt_first[0]
    sd.get_state_async(pin_out.data_ptr())
    sd.synchronize()
    pin_in.copy_(pin_out)

    def e2e_step():
        sd.set_state_async(pin_in.data_ptr())
        if dec is not None:
            solver.step(SSPRK2, dt_e2e)
            solver.finish()
        else:
            Solvers.SSPRK2(td, dt_e2e)
        sd.get_state_async(pin_in.data_ptr())

    e2e_step()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ee0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ee1.record()
    barrier()
    ms_e2e = ee0.elapsed_time(ee1)
    sd.synchronize()

    # ---- max over ranks ----
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_owned)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0].item()), float(t[1].item())
    cells_total = float(tot.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = cells_total * args.steps / (ms * 1e-3)
    e2e_value = cells_total * e2e_steps / (ms_e2e * 1e-3)
    peak, peak_src = peaks()
    # dominant kernel by measured time share
    tot_kms = sum(v[0] for v in ktimes.values()) or 1.0
    dom = max(ktimes, key=lambda k: ktimes[k][0])
    dom_ms, dom_cnt = ktimes[dom]
    dom_bytes = KERNEL_BYTES_PER_CELL[dom] * mesh.nt
    achieved = dom_bytes / (dom_ms / max(dom_cnt, 1) * 1e-3) / 1e9
    kern = {k: {"ms_per_launch": (v[0] / v[1] if v[1] else 0.0), "launches": v[1], "share": v[0] / tot_kms,
                "alg_GBps": (KERNEL_BYTES_PER_CELL[k] * mesh.nt / (v[0] / v[1] * 1e-3) / 1e9 if v[1] and v[0] > 0 else 0.0),
                "alg_bytes_per_cell": KERNEL_BYTES_PER_CELL[k]}
            for k, v in ktimes.items()}
    step_gbps = value / world * B_PER_CELL_UPDATE_SSPRK2 / 1e9
    # CPU baseline: the oracle, 1 thread (the reference is single-threaded), bounded sample
    cpu = None
    if not args.no_cpu:
        import __graft_entry__  # noqa: F401
        v_cpu, el_cpu, cells_cpu, thr = time_oracle(args.cpu_n, args.case, args.cpu_steps, 1, 1)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"StructTriangMesh({args.cpu_n},{args.cpu_n}) = {cells_cpu} cells of the same workload, "
                         f"{args.cpu_steps} SSPRK2 steps in {el_cpu:.1f} s, scalar oracle (upstream does not build here)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "wet_cell_fraction": wet_frac,
                   "cells_per_gpu": int(n_owned), "cells_total": int(cells_total),
                   "l2": "inputs larger than L2 (state + edge fields >> 126 MB), no flush needed",
                   "parallelism": "1 GPU" if world == 1 else (f"{world} strips, 3-row halo, {'peer-memory stores over NVLink (CUDA IPC)' if solver.halo.transport == 'p2p' else 'NCCL send/recv'} "
                                   f"{'overlapped with interior reconstruction' if solver.overlap else '(not overlapped)'} + min all-reduce"),
                   "dt": "CFLdt of the previous step (device resident)", "setup_s": t_setup,
                   "device_numbering": "hilbert" if args.reorder else "caller"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(24 * mesh.nt), "d2h_bytes_per_step": int(24 * mesh.nt),
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "what": "per step: swe_set_state_async(pinned host) + swe_step + swe_get_state_async(pinned host)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (NCU_TRAFFIC_BYTES_64M.get(dom) if (args.n == 4096 and args.case == "fully_wet") else None),
                     "traffic_source": "ncu --set full, profiles/r1_v3_ncu_full_summary.csv", "peak_source": peak_src,
                     "alg_bytes_per_launch": dom_bytes,
                     "step": {"alg_bytes_per_cell_update": B_PER_CELL_UPDATE_SSPRK2, "achieved": step_gbps,
                              "frac": step_gbps / peak, "per": "GPU"},
                     "kernels": kern},
        "cpu_baseline": cpu,
        "mass_drift_rel": ((d1["mass"] - d0["mass"]) / d0["mass"] if d0["mass"] else 0.0) if world == 1 else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=4096, help="squares per side per GPU (4 n^2 cells)")
    ap.add_argument("--global-n", type=int, default=0,
                    help="strong scaling: fixed global mesh of global_n x global_n squares split over the GPUs "
                         "(configs[4]: 8192 = 268M cells)")
    ap.add_argument("--case", default="fully_wet", choices=["fully_wet", "thacker"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-n", type=int, default=1024)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="swe_set_option key=value (A/B runs), repeatable")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: do not overlap the halo exchange with interior work")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1 halo transport: p2p = pack kernel stores into the peer GPU's buffer over NVLink "
                         "(CUDA IPC) + flag; nccl = pack, NCCL send/recv, unpack")
    ap.add_argument("--no-reorder", dest="reorder", action="store_false",
                    help="keep the caller's numbering on the device (default: Hilbert-curve renumbering of cells / "
                         "edges / nodes, A/B-measured +3.8 %% on this workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
