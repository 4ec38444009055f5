#!/bin/bash
# end-of-session check of the committed tree: raster / geometry / stack / icon GPU tests, a short bench line, and a bounded
# memcheck pass over the raster tests
bash tools/quick_check.sh
timeout 130 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_raster_gpu.py tests/test_edge_cases_gpu.py -q -x -k "not full_size and not bench_scene" 2>&1 | tail -3
