#!/bin/bash
# paths8k e2e at N GPUs under different geometry policies: N=$1
N=$1
nproc
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/probe_$tag.json 2> gpurun_out/probe_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/probe_$tag.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("$tag", "e2e", round(e["value"], 1), "ms", round(e["ms_per_step"], 1), "per rank", e.get("per_rank_ms"))
PY
}
run auto X=1
run device RB_GEO_MODE=1
run share10 RB_GEO_HOST_SHARE=0.10
run host RB_GEO_MODE=2
