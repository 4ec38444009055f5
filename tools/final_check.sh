#!/bin/bash
# What the driver runs at round end, plus the geometry-forced suite: GPU tests, smoke, default bench, reference arm, launch list.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
RB_GEO_MODE=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02c_bench_line.json 2> gpurun_out/r02c_bench.err; tail -2 gpurun_out/r02c_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_reference_line.json 2>> gpurun_out/r02c_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-table --no-configs > /dev/null 2>&1
python - <<PY
import json
for f in ("r02c_bench_line", "r02c_bench_reference_line"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", e.get("value"), e.get("ms_per_step"), "roof", (d.get("roofline") or {}).get("frac"), "parity", d.get("parity"), "launches", d.get("gpu_launches"))
    print("  geometry", json.dumps(e.get("geometry"))[:300])
    for k, c in (d.get("configs") or {}).items():
        print(" cfg", k, c.get("value"), c.get("ms_per_step"), "e2e", (c.get("e2e") or {}).get("value"), "parity", (c.get("parity") or {}).get("differing_px"))
PY
