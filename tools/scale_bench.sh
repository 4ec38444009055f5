#!/bin/bash
# bench.py at N GPUs of one box, launched as the driver does (run under gpurun --gpus N): N=$1
N=$1
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err
tail -2 gpurun_out/r02b_bench_n$N.err
RB_GEO_MODE=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02b_bench_n${N}_hostgeo.json 2>> gpurun_out/r02b_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --workload icons > gpurun_out/r02b_icons_n$N.json 2>> gpurun_out/r02b_bench_n$N.err
python - <<PY
import json
for f in ("r02b_bench_n$N", "r02b_bench_n${N}_hostgeo", "r02b_icons_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", json.dumps(d.get("e2e"))[:420])
    except Exception as e:
        print(f, "failed", e)
PY
