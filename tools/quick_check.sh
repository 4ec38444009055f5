#!/bin/bash
python -m pytest tests/test_geo_gpu.py tests/test_raster_gpu.py tests/test_icons.py tests/test_stack.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'prepass', round(d['roofline']['prepass_ms'],2), 'raster', round(d['roofline']['kernel_ms'],2), '| e2e', round(e['value'],1), round(e['ms_per_step'],1), 'record', round(e['record_ms'],1), 'build', round(e['host_build_and_enqueue_ms'],1), 'wait', round(e['gpu_wait_and_d2h_ms'],1))"
