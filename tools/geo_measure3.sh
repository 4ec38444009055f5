#!/bin/bash
python -m pytest tests/test_geo_gpu.py tests/test_raster_gpu.py -x -q 2>&1 | tail -3
nproc
for share in auto 0 0.2 0.3 0.4 0.5 0.6 1; do
  echo HOST_SHARE=$share
  if [ $share = auto ]; then python tools/geo_probe.py paths8k 5 2>&1 | tail -2; else RB_GEO_HOST_SHARE=$share python tools/geo_probe.py paths8k 5 2>&1 | tail -2; fi
done
echo MODE=2 host only
RB_GEO_MODE=2 python tools/geo_probe.py paths8k 5 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs > gpurun_out/bench_geo.json 2> gpurun_out/bench_geo.err; tail -3 gpurun_out/bench_geo.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_geo.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]))
PY
