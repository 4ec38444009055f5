#!/bin/bash
# ncu (stall reasons, instruction counts) of k_raster_warp's big launch with each named variant library (run under gpurun)
cp resvg_b200/libresvg_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp build_variants/$v.so resvg_b200/libresvg_b200.so
  ncu --clock-control none -k regex:k_raster_warp -c 1 --section WarpStateStats --section InstructionStats --section SchedulerStats --section LaunchStats --csv --log-file gpurun_out/ncu_variant_$v.csv \
     python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 1 > /dev/null 2>&1
done
cp /tmp/lib_keep.so resvg_b200/libresvg_b200.so
