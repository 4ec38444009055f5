#!/bin/bash
# banded download: parity tests, then the default bench's e2e leg with 1 / 4 / 6 / 8 / 12 bands and with the plain download
python -m pytest tests/test_geo_gpu.py tests/test_raster_gpu.py -x -q -m gpu 2>&1 | tail -2
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$1', 'value', round(d['value'],1), '| e2e', round(e['value'],1), round(e['ms_per_step'],2), 'build', round(e['host_build_and_enqueue_ms'],1), 'wait', round(e['gpu_wait_and_d2h_ms'],1))"
}
RB_BENCH_PLAIN_DOWNLOAD=1 run plain
for nb in 1 4 6 8 12; do RB_DL_BANDS=$nb run bands$nb; done
RB_BENCH_PLAIN_DOWNLOAD=1 run plain
