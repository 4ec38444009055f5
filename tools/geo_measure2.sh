#!/bin/bash
python -m pytest tests/test_geo_gpu.py -x -q 2>&1 | tail -3
for parts in 1 2 4 8; do
  echo PARTS=$parts
  RB_SUBMIT_PARTS=$parts python tools/geo_probe.py paths8k 5 2>&1 | tail -3
done
echo PARTS=4 own_stream=0
RB_GEO_OWN_STREAM=0 RB_SUBMIT_PARTS=4 python tools/geo_probe.py paths8k 5 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-configs > gpurun_out/bench_geo.json 2> gpurun_out/bench_geo.err; tail -3 gpurun_out/bench_geo.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_geo.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]))
PY
