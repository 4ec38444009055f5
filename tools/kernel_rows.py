"""Prints selected rows of the `kernels` table of a bench JSON line read from stdin: python bench.py ... | python tools/kernel_rows.py convolve morph"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
keys = sys.argv[1:]
for r in d.get("kernels", []):
    if not keys or any(k in r["kernel"] for k in keys):
        print(r)
