#!/bin/bash
python -m pytest tests/test_geo_gpu.py -x -q -m gpu -k submit_download 2>&1 | tail -2
RB_DL_DIAG=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 2 2>&1 >/dev/null | grep "dl diag" | tail -8
RB_DL_DIAG=1 RB_GEO_MODE=2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-table --no-configs --e2e-steps 2 2>&1 >/dev/null | grep "dl diag" | tail -8
