#!/usr/bin/env python
"""In-container only: render the reference's regression corpus (crates/resvg/tests/tests/**.svg) through the test-side
front end + CPU oracle and compare with the checked-in golden PNGs using the reference's own rule (±1 per demultiplied
channel, zero differing pixels, tests/integration/main.rs:151-226).  Prints a pass table; never used at test time."""
import collections
import glob
import os
import sys
import traceback

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import svgfront as F  # noqa: E402
from tests.backends import OracleBackend  # noqa: E402

CORPUS = "/root/reference/crates/resvg/tests/tests"


def run(patterns, verbose=False):
    be = OracleBackend()
    stats = collections.defaultdict(lambda: [0, 0, 0, 0])  # pass, fail, unsupported, error
    fails = []
    files = []
    for p in patterns:
        files += sorted(glob.glob(os.path.join(CORPUS, p), recursive=True))
    for svg in files:
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
            if n == 0:
                stats[fam][0] += 1
            else:
                stats[fam][1] += 1
                fails.append((rel, n))
        except (F.Unsupported, ImportError) as e:
            stats[fam][2] += 1
            if verbose:
                print("unsupported", rel, e)
        except Exception as e:  # front-end gap
            stats[fam][3] += 1
            if verbose:
                print("error", rel, repr(e))
                traceback.print_exc()
    tot = [0, 0, 0, 0]
    for fam in sorted(stats):
        s = stats[fam]
        print(f"{fam:45s} pass {s[0]:3d}  fail {s[1]:3d}  unsupported {s[2]:3d}  error {s[3]:3d}")
        tot = [a + b for a, b in zip(tot, s)]
    print(f"{'TOTAL':45s} pass {tot[0]:3d}  fail {tot[1]:3d}  unsupported {tot[2]:3d}  error {tot[3]:3d}")
    for rel, n in fails:
        print("FAIL", rel, n)
    return stats, fails


if __name__ == "__main__":
    pats = sys.argv[1:] or ["shapes/**/*.svg"]
    run(pats, verbose=os.environ.get("V") == "1")
