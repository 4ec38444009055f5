#!/bin/bash
python -m pytest tests/test_filters_gpu.py tests/test_icons.py -x -q -m gpu 2>&1 | tail -3
python -m pytest tests/test_golden.py -x -q -m gpu -k "GaussianBlur or DropShadow or blur or filter" 2>&1 | tail -2
for vh in 0 1; do
  echo RB_BOX_VH=$vh
  RB_BOX_VH=$vh python - <<PY
import resvg_b200 as rb, numpy as np
ctx = rb.Context(0)
W = H = 8192
a = ctx.layer(W, H)
rng = np.random.default_rng(1)
img = rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8); img[..., :3] = np.minimum(img[..., :3], img[..., 3:4])
big = np.tile(img, (8, 8, 1))
a.upload(big)
for sigma in (2.0, 4.0, 8.0, 9.0, 20.0):
    rb.filters.box_blur(sigma, sigma, a)
    ctx.timer_begin()
    for _ in range(5): rb.filters.box_blur(sigma, sigma, a)
    ms = ctx.timer_end() / 5
    print("sigma", sigma, "ms", round(ms, 3))
PY
done
python bench.py --workload filters8k --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('filters8k', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
