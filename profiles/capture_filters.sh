#!/bin/bash
# ncu passes over the filter chain (bench.py --workload filters8k) on one B200 (run through gpurun).
# Usage: profiles/capture_filters.sh TAG
#   gpurun_out/<tag>_filters_launches.csv : per-launch durations of the whole bench command
#   gpurun_out/<tag>_<kernel>.ncu-rep (+ _raw.csv): --set full of one launch each of the box-blur, convolve, morphology,
#   turbulence and lighting kernels
set -u
tag=${1:-r01}
mkdir -p gpurun_out
cmd="python bench.py --workload filters8k --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_filters_launches.csv \
    $cmd > gpurun_out/${tag}_filters_under_ncu.log 2>&1
for k in k_box_blur_h2 k_box_blur_v3 k_convolve_tile k_morph_tile k_turbulence k_lighting; do
    ncu --set full --clock-control none --import-source on -k regex:${k} -s 1 -c 1 -f \
        -o gpurun_out/${tag}_${k} $cmd > gpurun_out/${tag}_${k}_ncu_full.log 2>&1
    ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
done
