#!/usr/bin/env python
"""Writes the usvg trees of the `stack4k` workload (BASELINE configs[3]) as RBT1 streams: resvg_b200/data/stack4k_r<k>.rbt.

Parsing SVG into a usvg::Tree is host work that stays in Rust (usvg) and is out of scope here; in this repository the
test-side front end (tests/svgfront.py) plays usvg.  bench.py must not import tests/ on its product arm, so the trees of
the eight per-rank documents are produced once by this tool and committed — they are what the Rust shim would hand to
rb_tree_parse.  Deterministic: scenes.stack_svg(4096, 64, scene_seed(0x5EED0004, rank))."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resvg_b200 as rb  # noqa: E402
from resvg_b200 import scenes, shard  # noqa: E402
from tests import svgfront as F  # noqa: E402

SEED = 0x5EED0004


def main():
    out = os.path.join(ROOT, "resvg_b200", "data")
    os.makedirs(out, exist_ok=True)
    for rank in range(8):
        scene = F.parse(scenes.stack_svg(4096, 64, shard.scene_seed(SEED, rank)))
        blob = rb.tree.serialize(scene)
        rb.tree.Tree(blob).close()  # validates
        with open(os.path.join(out, f"stack4k_r{rank}.rbt"), "wb") as f:
            f.write(blob)
        print(rank, len(blob))
    # the parity case of bench.py: the same recipe at 1024 px
    scene = F.parse(scenes.stack_svg(1024, 64, SEED))
    with open(os.path.join(out, "stack1k.rbt"), "wb") as f:
        f.write(rb.tree.serialize(scene))


if __name__ == "__main__":
    main()
