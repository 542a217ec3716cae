#!/bin/bash
# A/B runs of tuning variants on one box (same process environment, back to back): prints ms/step and ms/stage per kernel.
B="python bench.py --no-cpu --no-repro --e2e-steps 2"
P='import json,sys; d=json.load(open(sys.argv[1])); k=d["roofline"]["kernels"]; print(sys.argv[1], round(d["ms_per_step"],3), {a:round(b["ms_per_stage"],3) for a,b in k.items()}, d["clocks"]["sm_mhz"], "thacker", d["thacker_basin"]["ms_per_step"] if d.get("thacker_basin") else None)'
run() { name=$1; shift; $B "$@" > gpurun_out/r2_ab_$name.json 2>gpurun_out/r2_ab_$name.err; python -c "$P" gpurun_out/r2_ab_$name.json || tail -3 gpurun_out/r2_ab_$name.err; }
run wet
run thacker_auto --case thacker --no-thacker
run thacker_off --case thacker --no-thacker --opt dry_skip=0
