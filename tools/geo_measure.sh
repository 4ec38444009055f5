#!/bin/bash
# Lane-layout sweep of the geometry kernels on the 100 000-path scene (run on the GPU box).
python -m pytest tests/test_geo_gpu.py -x -q 2>&1 | tail -3
for ls in ${LANES:-3}; do
  echo LANE_SHIFT=$ls
  RB_GEO_STREAMS=0 RB_GEO_LANE_SHIFT=$ls ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_geo -c 12 --csv --log-file gpurun_out/geo_launches_$ls.csv python tools/geo_probe.py paths8k 1 2>&1 | tail -1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/geo_launches_$ls.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
tot=0
for r in rows[1:]: print(r[ki][:24], r[gi], r[vi]); tot+=float(r[vi])
print("total ms", tot/1e6)
PY
  RB_GEO_LANE_SHIFT=$ls RB_GEO_DIAG=1 python tools/geo_probe.py paths8k 3 2>&1 | grep -o "wait [0-9.]* ms\|step.*" | tail -4
done
if [ -n "$NCU_FULL" ]; then
  RB_GEO_STREAMS=0 ncu --set full --clock-control none -k regex:"k_geo_stroke|k_geo_hair" -c 2 -o /tmp/geo_full -f python tools/geo_probe.py paths8k 1 2>&1 | tail -1
  ncu -i /tmp/geo_full.ncu-rep --page raw --csv > gpurun_out/geo_full_raw.csv; python tools/ncu_pick.py gpurun_out/geo_full_raw.csv
fi
