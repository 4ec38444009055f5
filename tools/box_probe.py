"""Timing of the box blur passes and a few other filter kernels on an 8192x8192 layer (CUDA events on the library's stream)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import resvg_b200 as rb

ctx = rb.Context(0)
W = H = 8192
rng = np.random.default_rng(1)
a = ctx.layer_from(rng.integers(0, 256, (H, W, 4), dtype=np.uint8))
F = rb.filters
line = []
for sx, sy in ((0.0, 4.0), (0.0, 20.0), (0.0, 64.0), (4.0, 0.0), (20.0, 0.0), (64.0, 0.0), (20.0, 20.0), (64.0, 64.0), (8.0, 8.0)):
    F.box_blur(sx, sy, a)
    ctx.timer_begin()
    for _ in range(3):
        F.box_blur(sx, sy, a)
    line.append(f"({sx:g},{sy:g}) {ctx.timer_end() / 3:.3f} ms")
print("box_blur  " + "  ".join(line))
for oct_ in (1, 2, 3, 4):
    F.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, oct_, 7, False, False, a)
    ctx.timer_begin()
    F.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, oct_, 7, False, False, a)
    print(f"turbulence {oct_} octaves: {ctx.timer_end():.3f} ms")
for r in (1.0, 3.0, 8.0, 32.0):
    F.morphology("dilate", r, r, a)
    ctx.timer_begin()
    F.morphology("dilate", r, r, a)
    print(f"morphology dilate r={r:g}: {ctx.timer_end():.3f} ms")
