"""Where does a stack4k document spend its time?  host enqueue (rb_render returns) vs device (CUDA events)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import resvg_b200 as rb
ctx = rb.Context(0)
blob = open("resvg_b200/data/stack4k_r0.rbt", "rb").read()
tree = rb.tree.Tree(blob)
ident = (1, 0, 0, 1, 0, 0)
target = ctx.layer(4096, 4096)
for _ in range(3):
    rb.tree.render(tree, ident, target)
ctx.synchronize()
rb._ffi.lib.rb_debug_profile(1)
for _ in range(3):
    l0 = ctx.launch_count
    ctx.timer_begin()
    t0 = time.perf_counter()
    rb.tree.render(tree, ident, target)
    t1 = time.perf_counter()
    ms = ctx.timer_end()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.2f} ms, device span {ms:.2f} ms, wall {1e3*(t2-t0):.2f} ms, launches {ctx.launch_count-l0}")
rb._ffi.lib.rb_debug_profile(0)
