#!/bin/bash
# paths8k at N GPUs of one box, launched as the driver does (run under gpurun --gpus N): N=$1
N=$1
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_n$N.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("N=$N value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(e["value"], 1), "ms", round(e["ms_per_step"], 1), e.get("per_rank_ms", {}).get("e2e"))
PY
