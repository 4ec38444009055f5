#!/bin/bash
# ncu --set full of k_raster_warp's big launch in the default bench (run under gpurun): raw metrics as CSV, per-line summary
ncu --set full --clock-control none --import-source on -k regex:k_raster_warp -c 1 -o /tmp/rw_full \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 1 > /dev/null 2>&1
ncu -i /tmp/rw_full.ncu-rep --page raw --csv > gpurun_out/r02c_k_raster_warp_ncu_full.csv 2>/dev/null
python tools/ncu_lines.py /tmp/rw_full.ncu-rep 600 > gpurun_out/r02c_k_raster_warp_lines_all.txt 2>&1
head -60 gpurun_out/r02c_k_raster_warp_lines_all.txt | cut -c1-160 > gpurun_out/r02c_k_raster_warp_source_lines.txt
python tools/phase_table.py gpurun_out/r02c_k_raster_warp_lines_all.txt > gpurun_out/r02c_k_raster_warp_phases.txt 2>&1
