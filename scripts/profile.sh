#!/bin/bash
# The profiling recipe behind profiles/ (run on a B200 box, e.g. via gpurun; see B200_PROFILING.md).
# Outputs go to gpurun_out/; copy the summaries you want to keep into profiles/.
set -e
mkdir -p gpurun_out
# 1. bench line (CUDA-event timing, clocks sampled during the timed region)
python bench.py > gpurun_out/bench.json
# 2. every launch with its device time (cold-cache, serialised: compare SHARES with the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
# 3. full capture of the four hot kernels (one launch each, after the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_flux|k_drain|k_update" -s 10 -c 5 \
    -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
# 4. read it (works without a GPU):
#    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|gpu__dram_throughput|pipe_fp64|registers_per_thread'
#    ncu -i gpurun_out/prof.ncu-rep --page source --csv --kernel-name regex:k_reconstruct   # stall reasons per SASS line
# 5. static checks: python swe_fvm_b200/build.py -v   (ptxas -v: registers / spills);  cuobjdump -sass swe_fvm_b200/libswe_b200.so
# 6. memory / race checks: compute-sanitizer --tool memcheck|racecheck|initcheck python scripts/sanitize_smoke.py
