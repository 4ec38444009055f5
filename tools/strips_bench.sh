#!/bin/bash
# stack4k in canvas strips at N GPUs: N=$1
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --workload stack4k --shard strips > gpurun_out/r02b_stack4k_strips_n$N.json 2> gpurun_out/r02b_stack4k_strips_n$N.err
tail -2 gpurun_out/r02b_stack4k_strips_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --shard strips > gpurun_out/r02b_paths8k_strips_n$N.json 2>> gpurun_out/r02b_stack4k_strips_n$N.err
python - <<PY
import json
for f in ("r02b_stack4k_strips_n$N", "r02b_paths8k_strips_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
