#!/bin/bash
# like variants_bench.sh, plus stack4k
cp resvg_b200/libresvg_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp build_variants/$v.so resvg_b200/libresvg_b200.so
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'raster', round(d['roofline']['kernel_ms'],2))"
  python bench.py --workload stack4k --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   stack4k', round(d['value'],1), round(d['ms_per_step'],2))"
done
cp /tmp/lib_keep.so resvg_b200/libresvg_b200.so
