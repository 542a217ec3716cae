#!/usr/bin/env python
"""bench.py — fp64 cell-updates/s of the SWE_FVM explicit time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n NI] [--case fully_wet|thacker]

A "step" is one SSPRK2 HLLC<Einfeldt> time step (2 stages) of every cell of the workload:
  N = 1 : configs[3] — synthetic StructTriangMesh(4096, 4096, 4/4096), 67 108 864 cells, fully-wet
          variant (the wet-cell fraction is printed), dt = CFLdt() of the previous step.
  N > 1 : weak scaling — rank r owns a 4096 x 4096 strip of StructTriangMesh(4096, 4096 N, h) plus
          3 halo rows per open side, all behind the swe_dist_* C-ABI: one peer-memory (NVLink) halo exchange
          per stage fused into the stage, one peer-memory min all-reduce per step. torch.distributed is only
          the bootstrap all-gather, the barrier around the timed region and the max-over-ranks.
          Before timing every run checks reproducibility ("repro"): 12 adaptive steps of a small Thacker basin
          on the same N-rank path against rank 0's single-context run (state hash + dt).
`value` = cells * K / (max over ranks of the CUDA-event time of K steps), state resident in HBM.
`e2e`   = same, through the host-buffer C-ABI calls: every step uploads the state from pinned
          host memory (swe_set_state_async), steps, and downloads it again (swe_get_state_async).
`--impl reference` times the CPU oracle (bit-identical to upstream's own sources compiled in oracle/_ref,
tests/test_ref_anchor.py) with all host threads on the SAME workload (mesh, case and state built by the oracle
itself, nothing of the product library is loaded). Prints ONE JSON line on rank 0.
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
    "k_flux": 150.0,         # per edge: slots 8 + n 16 + two sides 48 (+ dmin 8 on the LAST stage only: the CFL minimum of
                             # the earlier stages is a dead value) | F 24; x 1.5 edges per cell -> 144 (stage 1) / 156 (stage 2), mean 150
    "k_drain": 56.0,         # te 12 + F0 12 + w 8 + cb 8 + area 8 | dti 8
    "k_halo_pack_signal": 0.0, "k_halo_wait_unpack": 0.0, "k_min_push_pull": 0.0,  # O(sqrt N) halo / scalar kernels (N > 1)
    "k_update": 220.0,       # W 24 + (U0 24 on the 2nd stage) + te,tt 24 + F 36 + dti 8 + ceh 24 + grad 16 + n 24 +
                             # L 12 + area 8 + cb 8 | W 24  -> 208 (stage 1) / 232 (stage 2), mean 220
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# profiles/r2_ncu_full_summary.csv (64M-cell fully-wet workload); None for other workloads
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
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "timestamp")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self, which: str):
        """Wall-clock bounds of the timed region: the sampler is started BEFORE the warm-up (nvidia-smi needs a few hundred
        milliseconds to come up, longer than a short timed region) and only the samples stamped inside the region count."""
        setattr(self, "t_" + which, time.time())

    @staticmethod
    def _stamp(txt: str):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

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
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                rows.append((self._stamp(c[9]), float(c[1]), float(c[2]), float(c[3]), c[5:9]))
            except ValueError:
                continue
        t0, t1 = getattr(self, "t_begin", None), getattr(self, "t_end", None)
        inside = [r for r in rows if r[0] is not None and t0 is not None and t1 is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        if not inside and rows and t0 is not None and t1 is not None:  # region shorter than the sampling period: nearest samples
            mid = 0.5 * (t0 + t1)
            near = sorted((r for r in rows if r[0] is not None), key=lambda r: abs(r[0] - mid))[:2]
            inside = [r for r in near if abs(r[0] - mid) <= 0.5 * (t1 - t0) + 0.25]
        if t0 is None:
            inside = rows
        sm, mx, pw, reasons = [], [], [], set()
        for _, a, b, p_, flags in inside:
            sm.append(a); mx.append(b); pw.append(p_)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
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


def make_case_obj(kind: str, mid_y: float, length: float):
    from swe_fvm_b200 import Case
    return Case("classic_thacker" if kind == "thacker" else "fully_wet", 2.0, mid_y, length)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle alone (cpu_baseline of the GPU line, and --impl reference). Mesh, bathymetry and
# initial state come from the oracle's own generator / cases: no product code is loaded.
# ----------------------------------------------------------------------------------------------
def time_oracle(n: int, case_kind: str, steps: int, warmup: int, threads: int):
    from oracle.oracle import Oracle, OracleCase, OracleStructMesh
    t0 = time.perf_counter()
    mesh = OracleStructMesh(n, n, 4.0 / n)
    case = OracleCase("classic_thacker" if case_kind == "thacker" else "fully_wet", 2.0, 2.0, 4.0)
    case.set_bathymetry(mesh)
    v0 = case.initial_state(mesh, quad_n=1)
    o = Oracle(mesh, threads=threads)
    o.set_state(v0)
    del v0
    setup = time.perf_counter() - t0
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
    return mesh.nt * steps / el, el, mesh.nt, o.threads, setup


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as o
    o.build()
    threads = os.cpu_count() or 1
    # the full configs[3] mesh needs ~45 GB of host memory in the oracle's int64 / AoS layout
    try:
        import psutil
        free_gb = psutil.virtual_memory().available / 2**30
    except Exception:
        free_gb = 0.0
    n = args.n if free_gb >= 64.0 else min(args.n, 2048)
    rate1, _, _, _, _ = time_oracle(128, args.case, 2, 1, threads)
    # keep the whole run within a few minutes: bound the number of timed steps, not the mesh
    per_step = 4.0 * n * n / max(rate1, 1.0)
    steps = max(2, min(args.steps, int(150.0 / max(per_step, 1e-9))))
    warmup = 1 if per_step > 2.0 else min(args.warmup, 3)
    v, el, cells, thr, setup = time_oracle(n, args.case, steps, warmup, threads)
    same = (n == args.n)
    sample = (f"StructTriangMesh({n},{n},4/{n}) = {cells} cells of the {args.case} workload, {steps} timed SSPRK2 "
              f"HLLC<Einfeldt> steps (asked: {args.steps}) after {warmup} warm-up, dt=CFLdt, OpenMP oracle (snapshot semantics), "
              f"mesh + state built by the oracle in {setup:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, max(args.gpus, 1)), "sample": sample, "same_config": same,
                   "host": {"nproc": os.cpu_count(), "cpu_model": cpu_model(), "free_gb": round(free_gb, 1)}},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the oracle is a scalar C++ restatement of upstream's path, pinned bit for bit to upstream's own src/*.cpp "
                "compiled in oracle/_ref (Eigen-subset shim; too slow to time: it recomputes the geometry per access like "
                "upstream); OpenMP over the snapshot-semantics loops",
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
def repro_check(args, rank, world, local_rank):
    """Driver-visible correctness of the N-rank path: 12 adaptive SSPRK2 steps of a Thacker basin whose front crosses
    the strip boundaries, through the SAME swe_dist path that is timed below (peer memory, overlap on), against rank
    0's single-context run of the undecomposed mesh: order-independent state hash (equal iff every cell is
    bit-identical) and the final dt."""
    import torch
    import torch.distributed as dist
    from swe_fvm_b200 import Case, StructTriangMesh
    from swe_fvm_b200 import dist as swd
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc
    ni, nj, steps = 256, 64 * max(world, 2), 12
    h = 4.0 / ni
    case = Case("classic_thacker", 2.0, 0.5 * nj * h, 4.0)
    if world > 1:
        plan = swd.Plan.struct(rank, world, ni, nj, h)
        ds = swd.DistSolver(plan, device=local_rank, reorder=args.reorder, overlap=not args.no_overlap)
        ds.sd.set_case_bathymetry(case)
        ds.sd.set_case_state(case, quad_n=2)
        ds.exchange()
        ds.run("ssprk2", steps, dt=0.0, dt0=1e-3)
        ds.synchronize()
        part = ds.state_hash()
        parts = [None] * world
        dist.all_gather_object(parts, part)
        got_hash = sum(parts) % (1 << 64)
        got_dt = ds.cfl_dt()
        ds.close()
    if rank != 0:
        return None
    mesh = StructTriangMesh(ni, nj, h)
    sd = SpaceDisc("hllc", "einfeldt", mesh, None, device=local_rank, reorder=args.reorder)
    sd.set_case_bathymetry(case)
    sd.set_case_state(case, quad_n=2)
    m0 = sd.diagnostics()["mass"]
    td = TimeDisc(sd)
    Solvers.run(td, "ssprk2", steps, dt=0.0, dt0=1e-3)
    sd.synchronize()
    want_hash, want_dt = sd.state_hash(), td.CFLdt()
    d1 = sd.diagnostics()
    out = {"n_cells": mesh.nt, "steps": steps, "workload": f"ClassicThacker on StructTriangMesh({ni},{nj}), adaptive SSPRK2 HLLC<Einfeldt>",
           "wet_cells": d1["wet_cells"], "mass_drift_rel": (d1["mass"] - m0) / m0}
    if world > 1:
        out.update({"hash_equal": got_hash == want_hash, "dt_equal": got_dt == want_dt, "ranks": world,
                    "what": "sum of the ranks' owned-cell state hashes == hash of the single-context run; final CFLdt identical"})
    else:
        out.update({"hash_equal": True, "dt_equal": True, "ranks": 1, "what": "single context (nothing to compare at N = 1)"})
    sd.close()
    return out


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
        # NCCL announces its version on stdout when the communicator is created (NCCL_DEBUG=VERSION on some boxes):
        # keep stdout for the one JSON line, send everything else to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if rank == 0:
        import __graft_entry__ as g
        g.build()
    if world > 1:
        dist.barrier()
    from swe_fvm_b200 import StructTriangMesh
    from swe_fvm_b200 import dist as swd
    from swe_fvm_b200.solver import Solvers, SpaceDisc, TimeDisc

    repro = repro_check(args, rank, world, local_rank) if not args.no_repro else None

    n, h = args.n, 4.0 / args.n
    t_setup = time.perf_counter()
    ds = None
    if args.global_n:  # strong scaling: one fixed global mesh, split into `world` strips
        n, h = args.global_n, 4.0 / args.global_n
        ni, nj = n, n
    else:
        ni, nj = n, n * world
    length_y = nj * h
    if world == 1:
        mesh = StructTriangMesh(ni, nj, h)
        t_mesh = time.perf_counter() - t_setup
        sd = SpaceDisc("hllc", "einfeldt", mesh, None, device=local_rank, reorder=args.reorder)
        n_owned = mesh.nt
    else:
        plan = swd.Plan.struct(rank, world, ni, nj, h)
        t_mesh = time.perf_counter() - t_setup
        mesh, n_owned = plan.mesh, plan.n_owned
        ds = swd.DistSolver(plan, device=local_rank, reorder=args.reorder, overlap=not args.no_overlap)
        sd = ds.sd
    t_create = time.perf_counter() - t_setup - t_mesh
    sd.set_stream(torch.cuda.current_stream().cuda_stream)
    for kv in args.opt:
        k, v = kv.split("=")
        sd.set_option(k, int(v))
    td = TimeDisc(sd)
    nt_local = mesh.nt

    def init_case(kind):
        case = make_case_obj(kind, 0.5 * length_y, 4.0)
        sd.set_case_bathymetry(case)          # nodal bed and cell-average initial state on the device
        sd.set_case_state(case, quad_n=1)
        if ds is not None:
            ds.exchange()
        sd.synchronize()

    init_case(args.case)
    d0 = sd.diagnostics()
    wet_frac = d0["wet_cells"] / nt_local
    t_setup = time.perf_counter() - t_setup

    dt_first = [1e-6]

    def run_steps(k):
        if ds is not None:
            ds.run("ssprk2", k, dt=0.0, dt0=dt_first[0])
        else:
            Solvers.run(td, "ssprk2", k, dt=0.0, dt0=dt_first[0])

    def sync():
        if ds is not None:
            ds.synchronize()
        else:
            sd.synchronize()

    def cfl():
        return ds.cfl_dt() if ds is not None else td.CFLdt()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        run_steps(k)
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1)

    # prime the CFL dt with one tiny step, then warm up (the clock sampler starts here so that it is up in time)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    run_steps(1); sync()
    dt_first[0] = cfl()
    run_steps(max(args.warmup, 1)); sync()
    dt_first[0] = cfl()

    # ---- timed region: K steps, state resident in HBM ----
    sd.kernel_timing(True)
    launches0 = sd.launch_count()
    if sampler:
        sampler.mark("begin")
    ms = timed(args.steps)
    if sampler:
        sampler.mark("end")
    clocks = sampler.stop() if sampler else None
    launches = sd.launch_count() - launches0
    ktimes = sd.kernel_times()
    sd.kernel_timing(False)
    sync()  # raises if a non-finite state was produced or a peer wait timed out
    d1 = sd.diagnostics()

    # ---- e2e: host buffers in and out every step (C-ABI with pinned host memory) ----
    # swe_submit_step_host: every step uploads ITS input state from pinned host memory, runs one SSPRK2 step and
    # downloads the result into pinned host memory; consecutive steps are independent batches (two alternating
    # input and output buffers), so the library overlaps the upload of step n+1 and the download of step n-1 with
    # the compute of step n (PCIe is full duplex).
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    dt_e2e = cfl()
    pins_in = [torch.empty((nt_local, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
    pins_out = [torch.empty((nt_local, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
    sd.get_state_async(pins_in[0].data_ptr())
    sync()
    pins_in[1].copy_(pins_in[0])
    submit = ds.submit_step_host if ds is not None else sd.submit_step_host
    wait_host = ds.wait_host if ds is not None else sd.wait_host

    def e2e_run(k):
        for q in range(k):
            submit(pins_in[q & 1].data_ptr(), pins_out[q & 1].data_ptr(), "ssprk2", dt_e2e)
        wait_host()

    e2e_run(2)
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ee0.record()
    e2e_run(e2e_steps)          # returns after the last result has landed in host memory
    ee1.record()
    barrier()
    ms_e2e = ee0.elapsed_time(ee1)
    e2e_ok = bool(torch.isfinite(pins_out[0]).all().item() and torch.isfinite(pins_out[1]).all().item())
    # the same, strictly serial (one buffer, upload -> step -> download on one stream), for comparison
    sd.set_state_async(pins_in[0].data_ptr()); sync()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    es0.record()
    for _ in range(3):
        sd.set_state_async(pins_in[0].data_ptr())
        if ds is not None:
            ds.step("ssprk2", dt_e2e); ds.synchronize()
        else:
            Solvers.SSPRK2(td, dt_e2e)
        sd.get_state_async(pins_out[0].data_ptr())
    es1.record()
    barrier()
    ms_e2e_serial = es0.elapsed_time(es1) / 3.0
    sync()

    # ---- second sub-record at N = 1: the Thacker basin itself (configs[3] as named; mostly dry) ----
    thacker = None
    if world == 1 and args.case == "fully_wet" and not args.no_thacker:
        init_case("thacker")
        dt_first[0] = 1e-6
        run_steps(1); sync()
        dt_first[0] = cfl()
        run_steps(3); sync()
        dt_first[0] = cfl()
        k = max(5, args.steps // 2)
        ms_t = timed(k)
        sync()
        dt = sd.diagnostics()
        thacker = {"workload": f"configs[3] as named: ClassicThacker basin on StructTriangMesh({n},{n}), HLLC<Einfeldt>, SSPRK2, dt=CFLdt",
                   "value": nt_local * k / (ms_t * 1e-3), "unit": UNIT, "steps": k, "ms_per_step": ms_t / k,
                   "wet_cell_fraction": dt["wet_cells"] / nt_local}

    # ---- max over ranks ----
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_owned), float(nt_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0].item()), float(t[1].item())
    cells_total, local_total = float(tot[0].item()), float(tot[1].item())
    if rank != 0:
        if ds is not None:
            ds.close()
        if world > 1:
            dist.destroy_process_group()
        return

    value = cells_total * args.steps / (ms * 1e-3)
    e2e_value = cells_total * e2e_steps / (ms_e2e * 1e-3)
    peak, peak_src = peaks()
    # per-kernel numbers per STAGE (K1 / K4 are issued as several range launches per stage on N > 1 ranks):
    # stages = launches of the flux kernel, one per stage
    stages = max(ktimes.get("k_flux", (0.0, 0))[1], 1)
    tot_kms = sum(v[0] for v in ktimes.values()) or 1.0
    dom = max((k for k in ktimes if KERNEL_BYTES_PER_CELL.get(k, 0.0) > 0.0), key=lambda k: ktimes[k][0])
    dom_ms_stage = ktimes[dom][0] / stages
    dom_bytes = KERNEL_BYTES_PER_CELL[dom] * nt_local
    achieved = dom_bytes / (dom_ms_stage * 1e-3) / 1e9
    kern = {k: {"ms_per_stage": v[0] / stages, "launches": v[1], "share": v[0] / tot_kms,
                "alg_GBps": (KERNEL_BYTES_PER_CELL[k] * nt_local / (v[0] / stages * 1e-3) / 1e9 if v[0] > 0 else 0.0),
                "alg_bytes_per_cell": KERNEL_BYTES_PER_CELL[k]}
            for k, v in ktimes.items() if v[1] > 0}
    kernel_ms_per_step = 2.0 * sum(v["ms_per_stage"] for v in kern.values())
    own_bytes_stage = sum(KERNEL_BYTES_PER_CELL.values())
    step_gbps_contract = value / world * B_PER_CELL_UPDATE_SSPRK2 / 1e9
    step_gbps_own = value / world * 2.0 * own_bytes_stage / 1e9
    # CPU baseline: the oracle, 1 thread (the reference is single-threaded), bounded sample
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as _o
        _o.build()
        v_cpu, el_cpu, cells_cpu, thr, _ = time_oracle(args.cpu_n, args.case, args.cpu_steps, 1, 1)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": thr, "kind": "port", "cpu_model": cpu_model(), "nproc": os.cpu_count(),
               "sample": f"StructTriangMesh({args.cpu_n},{args.cpu_n}) = {cells_cpu} cells of the same workload, "
                         f"{args.cpu_steps} SSPRK2 steps in {el_cpu:.1f} s, scalar oracle (bit-identical to upstream's sources in oracle/_ref)"}
    traffic_ok = (args.n == 4096 and args.case == "fully_wet" and not args.global_n and world == 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "wet_cell_fraction": wet_frac,
                   "cells_per_gpu": int(n_owned), "cells_total": int(cells_total),
                   "l2": "inputs larger than L2 (state + edge fields >> 126 MB), no flush needed",
                   "parallelism": "1 GPU" if world == 1 else (
                       f"{world} strips behind the swe_dist_* C-ABI, 3-row halo, peer-memory stores over NVLink (CUDA IPC) fused into the "
                       f"stage ({'boundary cells updated first, stores overlap the interior update + reconstruction' if not args.no_overlap else 'not overlapped'}), "
                       f"peer-memory min all-reduce"),
                   "dt": "CFLdt of the previous step (device resident)", "setup_s": t_setup, "setup_breakdown_s": {"host_mesh": t_mesh, "create_context": t_create, "case_on_device": t_setup - t_mesh - t_create},
                   "initial_state": "device-side TriangAverage<3,1> of the analytic case",
                   "device_numbering": "hilbert" if args.reorder else "caller"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(24 * local_total), "d2h_bytes_per_step": int(24 * local_total),
                "bytes_are": "total over all ranks (local state incl. halo cells)", "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "results_finite": e2e_ok, "ms_per_step_serial": ms_e2e_serial,
                "what": "per step: swe_submit_step_host / swe_dist_submit_step_host = H2D of the step's input from pinned host memory + one "
                        "SSPRK2 step + D2H of the result into pinned host memory; independent consecutive steps are pipelined by the library "
                        "(upload n+1 / step n / download n-1 overlap); ms_per_step_serial = the same without overlap"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (NCU_TRAFFIC_BYTES_64M.get(dom) if traffic_ok else None),
                     "traffic_source": "ncu --set full, profiles/r2_ncu_full_summary.csv", "peak_source": peak_src,
                     "alg_bytes_per_launch": dom_bytes, "per": "stage (sum of the kernel's range launches), rank 0",
                     "step": {"alg_bytes_per_cell_update_contract": B_PER_CELL_UPDATE_SSPRK2, "achieved_contract": step_gbps_contract,
                              "frac_contract": step_gbps_contract / peak,
                              "alg_bytes_per_cell_update_this_layout": 2.0 * own_bytes_stage, "achieved": step_gbps_own,
                              "frac": step_gbps_own / peak, "per": "GPU",
                              "note": "frac uses the bytes THIS layout moves (sum of the kernels' algorithmic bytes); "
                                      "frac_contract uses SURVEY App. D's 744 B/cell-stage decomposition, which this layout undercuts"},
                     "kernels": kern, "kernel_ms_per_step": kernel_ms_per_step,
                     "non_kernel_ms_per_step": ms / args.steps - kernel_ms_per_step},
        "cpu_baseline": cpu,
        "repro": repro,
        "mass_drift_rel": (d1["mass"] - d0["mass"]) / d0["mass"] if d0["mass"] else 0.0,
        "mass_drift_of": "rank 0's local cells (owned + halo)" if world > 1 else "all cells",
        "thacker_basin": thacker,
    }
    print(json.dumps(line), flush=True)
    if ds is not None:
        ds.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=4096, help="squares per side per GPU (4 n^2 cells)")
    ap.add_argument("--global-n", type=int, default=0,
                    help="strong scaling: fixed global mesh of global_n x global_n squares split over the GPUs "
                         "(configs[4]: 8192 = 268M cells)")
    ap.add_argument("--case", default="fully_wet", choices=["fully_wet", "thacker"])
    ap.add_argument("--e2e-steps", type=int, default=20, help="steps of the host-buffer pipeline (its fill and drain are inside the timed region)")
    ap.add_argument("--cpu-n", type=int, default=1024)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="swe_set_option key=value (A/B runs), repeatable")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: do not overlap the halo exchange with interior work")
    ap.add_argument("--no-repro", action="store_true", help="skip the reproducibility check before the timed region")
    ap.add_argument("--no-thacker", action="store_true", help="N = 1: skip the second sub-record (the Thacker basin itself)")
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
