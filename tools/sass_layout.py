#!/usr/bin/env python
"""Code layout of one kernel: instructions per source region, from `nvdisasm -g -c` of a cubin built with -lineinfo.
usage: sass_layout.py <object or cubin> <kernel name substring> [bucket]"""
import collections, os, re, subprocess, sys, tempfile

def disasm(path):
    if path.endswith(".cubin"):
        cub = path
    else:
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cub = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
    return subprocess.run(["nvdisasm", "-g", "-c", cub], check=True, capture_output=True, text=True).stdout.split("\n")

def main():
    lines = disasm(sys.argv[1]); key = sys.argv[2]; bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and key in l][0]
    end = [i for i, l in enumerate(lines) if l.startswith("//--------------------- .text") and i > start]
    end = end[0] if end else len(lines)
    cur = None; n = 0; per = collections.Counter(); first = {}
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)) // bucket * bucket); continue
        if re.search(r"/\*[0-9a-f]{4,6}\*/", l) and ";" in l:
            n += 1; per[cur] += 1; first.setdefault(cur, n)
    print("instructions", n, "bytes", n * 16)
    for k in sorted(per, key=lambda k: (k[0], k[1]) if k else ("", 0)):
        print(f"{k[0] if k else '?':28s} {k[1] if k else 0:5d}  {per[k]:5d} instr {per[k]*16/1024:6.2f} KB  first at {first[k]*16/1024:6.1f} KB")

main()
