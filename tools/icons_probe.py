"""Quick timing probe of the icon atlas renderer (one 32x32 atlas = 1024 documents)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import resvg_b200 as rb
from resvg_b200 import documents

ctx = rb.Context(0)
atlas = documents.IconAtlas(ctx)
t0 = time.perf_counter(); sc = documents.prepare_chunk(0, 1024); t1 = time.perf_counter()
print("generate %.1f ms, %d paths" % ((t1 - t0) * 1e3, sc["n_paths"]))
for rep in range(3):
    t0 = time.perf_counter(); ch = atlas.render(sc); ctx.synchronize(); t1 = time.perf_counter()
    print("render (submit) %.2f ms" % ((t1 - t0) * 1e3), ch["base"].stats())
    atlas.release(ch)
ch = atlas.prepare(sc); ctx.synchronize()
for rep in range(3):
    ctx.timer_begin(); atlas.run(ch); ms = ctx.timer_end()
    t0 = time.perf_counter(); atlas.run(ch); ctx.synchronize(); t1 = time.perf_counter()
    print("run resident: device %.2f ms, wall %.2f ms; last batch prepass/raster ms %s" % (ms, (t1 - t0) * 1e3, ctx.last_run_ms()))
ns = dict(ch, shadow_docs=ch["shadow_docs"][:0])
ctx.timer_begin(); atlas.run(ns); print("without shadow docs: %.2f ms" % ctx.timer_end())
