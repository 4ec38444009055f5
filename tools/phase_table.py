#!/usr/bin/env python
"""Per-phase executed instructions of k_raster_warp from a tools/ncu_lines.py listing: phase_table.py LISTING PAIRS"""
import re, collections, sys
src = open('resvg_b200/csrc/raster_warp.cuh').read().split('\n')
def find(s, after=0):
    for i, l in enumerate(src):
        if i >= after and s in l: return i + 1
    raise KeyError(s)
marks = [("blend_tile_gradient", find("void blend_tile_gradient(")), ("setup / load dst", find("struct WarpDraw")), ("group prep", find("---- every lane prepares")), ("pair header", find("for (int k = 0; k < n_group")),
         ("hair", find("---- hairline stroke:")), ("pair bounds", find("const uint32_t bounds = __shfl_sync")), ("scatter", find("---- scatter")),
         ("backdrop", find("winding every sub-scanline starts")), ("scan", find("---- scan:")), ("coverage", find("---- coverage:")),
         ("clear", find("---- clear the marks")), ("blend (inline)", find("---- blend")), ("store / wrapper", find("    if (px_stats) {", find("---- blend")))]
rsrc = open('resvg_b200/csrc/raster.cu').read().split('\n')
def rfind(s):
    for i, l in enumerate(rsrc):
        if s in l: return i + 1
    raise KeyError(s)
def rfind_opt(s):
    try: return rfind(s)
    except KeyError: return None
rmarks = [m for m in [("R helpers", 1), ("R gradient t", rfind_opt("float gradient_t_at(")), ("R gradient colour", rfind_opt("PF gradient_color_at(")),
          ("R pattern", rfind_opt("float ulp_sub(")), ("R per-pixel gradient fns", rfind_opt("P16 shade16_gradient(")),
          ("R blend_pixel", rfind_opt("uint32_t blend_pixel(")), ("R rest", rfind("uint32_t blend_pixel(") + 60)] if m[1]]
def phase(marks, ln):
    name = marks[0][0]
    for n, a in marks:
        if ln >= a: name = n
    return name
tot = collections.Counter(); samp = collections.Counter()
for l in open(sys.argv[1]):
    m = re.match(r'(\S+):\s*(\d+)\s+([\d.]+)M\s+[\d.]+% samp\s+(\d+)', l)
    if not m: continue
    f, ln, n, s = m.group(1), int(m.group(2)), float(m.group(3)), int(m.group(4))
    if f == 'raster_warp.cuh': key = phase(marks, ln) if ln >= marks[0][1] else "k_row_lists / exact_span_break_list etc."
    elif f == 'raster.cu': key = phase(rmarks, ln)
    else: key = f
    tot[key] += n; samp[key] += s
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 7.69e6
T = sum(tot.values()); S = sum(samp.values())
print(f"total {T/1e3:.2f} G instructions, {T*1e6/pairs:.0f} per (draw, tile) pair")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v > 1: print(f"{k:28s} {v:9.1f}M {100*v/T:5.1f}%  per pair {v*1e6/pairs:6.0f}  stall samples {100*samp[k]/S:5.1f}%")
