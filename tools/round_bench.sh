#!/bin/bash
# The round's measured records (run on the GPU box): default bench line, reference arm, launch list of the same command.
python bench.py > gpurun_out/r02b_bench_line.json 2> gpurun_out/r02b_bench.err; tail -2 gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_line.json 2>> gpurun_out/r02b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-table --no-configs > /dev/null 2>&1
python - <<PY
import json
for f in ("r02b_bench_line", "r02b_bench_reference_line"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d.get("value"), d.get("ms_per_step"), "e2e", json.dumps(d.get("e2e"))[:600], "roof", (d.get("roofline") or {}).get("frac"), "parity", d.get("parity"))
    for k, c in (d.get("configs") or {}).items():
        print(" cfg", k, c.get("value"), c.get("ms_per_step"), "e2e", (c.get("e2e") or {}).get("value"), "parity", c.get("parity"))
PY
