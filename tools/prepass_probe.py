"""Prints the device time of the binning / edge-list pre-pass and of the raster kernel for the paths8k scene."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-kernel-table", "--steps", "5", "--e2e-steps", "1"], capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print("step %.2f ms  prepass %.3f ms  raster %.2f ms" % (d["ms_per_step"], d["roofline"]["prepass_ms"], d["roofline"]["kernel_ms"]))
