#!/bin/bash
python tools/geo_bisect_full.py 2>&1 | cut -c1-300 | head -8
python -m pytest tests/test_geo_gpu.py tests/test_raster_gpu.py -x -q 2>&1 | tail -3
RB_GEO_MODE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
