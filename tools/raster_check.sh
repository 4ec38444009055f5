#!/bin/bash
# raster kernel iteration: parity tests that exercise k_raster_warp, then the default bench (device-timed value, raster diag)
python -m pytest tests/test_raster_gpu.py tests/test_edge_cases_gpu.py tests/test_stack.py tests/test_icons.py tests/test_geo_gpu.py -x -q -m gpu 2>&1 | tail -2
RB_RASTER_DIAG=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs 2> gpurun_out/raster_check.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'prepass', round(d['roofline']['prepass_ms'],2), 'raster', round(d['roofline']['kernel_ms'],2), '| e2e', round(e['value'],1), round(e['ms_per_step'],1))"
grep "raster diag" gpurun_out/raster_check.err | head -3
python bench.py --workload stack4k --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stack4k', round(d['value'],1), round(d['ms_per_step'],2))"
