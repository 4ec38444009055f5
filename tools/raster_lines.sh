#!/bin/bash
# per-source-line executed instructions / stall samples of k_raster_warp's big launch (variant $1), via ncu SourceCounters
v=$1
cp resvg_b200/libresvg_b200.so /tmp/lib_keep.so
cp build_variants/$v.so resvg_b200/libresvg_b200.so
ncu --clock-control none -k regex:k_raster_warp -c 1 --section SourceCounters --section InstructionStats --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --import-source on -o /tmp/rw_$v \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 1 > /dev/null 2>&1
python tools/ncu_lines.py /tmp/rw_$v.ncu-rep 600 > gpurun_out/raster_lines_$v.txt 2>&1
cp /tmp/lib_keep.so resvg_b200/libresvg_b200.so
