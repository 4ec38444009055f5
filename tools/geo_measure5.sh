#!/bin/bash
RB_GEO_SUBRANGES=3 python -m pytest tests/test_geo_gpu.py -x -q 2>&1 | tail -3
for subs in 1 2 3 4; do
  echo threads=4 SUBRANGES=$subs
  RB_GEO_SUBRANGES=$subs python tools/geo_probe.py paths8k 5 0 4 2>&1 | tail -2
done
echo threads=16 default
python tools/geo_probe.py paths8k 5 0 16 2>&1 | tail -2
echo threads=16 SUBRANGES=2
RB_GEO_SUBRANGES=2 python tools/geo_probe.py paths8k 5 0 16 2>&1 | tail -2
