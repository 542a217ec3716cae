#!/bin/bash
# The profiling recipe behind profiles/ (run on a B200 box, e.g. via gpurun; see B200_PROFILING.md).
# Outputs go to gpurun_out/; scripts/summarise_profiles.py turns them into the committed profiles/rN_* files.
set -e
R=${1:-r2}
mkdir -p gpurun_out
# 1. bench line (CUDA-event timing, clocks sampled during the timed region) + the reference arm on the host cores
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err || true
# 2. every launch with its device time (cold-cache, serialised: compare SHARES with the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-thacker --no-repro --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
# 3. full capture of the four hot kernels (one launch each, after the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:"k_reconstruct|k_flux|k_drain|k_update" -s 24 -c 11 \
    -o gpurun_out/${R}_prof python bench.py --steps 2 --warmup 3 --no-cpu --no-thacker --no-repro --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
# 4. read it (works without a GPU):
#    ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv | python scripts/summarise_profiles.py ...
# 5. static checks: python swe_fvm_b200/build.py -v   (ptxas -v: registers / spills);  cuobjdump -sass swe_fvm_b200/libswe_b200.so
# 6. memory / race checks: compute-sanitizer --tool memcheck|racecheck|initcheck python scripts/sanitize_smoke.py
