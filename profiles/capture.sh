#!/bin/bash
# Profiling pass of the bench command on one B200 (run through gpurun).  Usage: profiles/capture.sh TAG [KERNEL_REGEX]
# Writes gpurun_out/<tag>_launches.csv (per-launch durations of the bench) and gpurun_out/<tag>_<kernel>.ncu-rep
# (+ raw csv) for one launch of the dominant kernel.
set -u
tag=${1:-r01}
kern=${2:-k_raster_warp}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${kern} -s 3 -c 1 -f \
    -o gpurun_out/${tag}_${kern} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_${kern}.ncu-rep --page raw --csv > gpurun_out/${tag}_${kern}_raw.csv 2>/dev/null
