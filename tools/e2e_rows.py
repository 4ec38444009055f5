import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print(round(e["ms_per_step"],1), round(e["host_build_and_enqueue_ms"],1), round(e["gpu_wait_and_d2h_ms"],1))
