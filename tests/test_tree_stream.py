"""The RBT1 tree stream (include/resvg_b200.h, "Whole-tree rendering"): the writer of resvg_b200/tree.py against the
library's parser.  Host-only — rb_tree_parse / rb_tree_size / rb_tree_node_bbox do no device work."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

import resvg_b200 as rb
from resvg_b200._ffi import lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _scene(name):
    with open(os.path.join(HERE, "golden", "scenes", name + ".json")) as f:
        return json.load(f)


def _parse(blob):
    h = C.c_void_p()
    st = lib.rb_tree_parse(blob, len(blob), C.byref(h))
    if st == 0:
        lib.rb_tree_destroy(h)
    return st


def test_every_fixture_scene_serialises_and_parses():
    n = 0
    for path in sorted(glob.glob(os.path.join(HERE, "golden", "scenes", "*.json"))):
        with open(path) as f:
            scene = json.load(f)
        blob = rb.tree.serialize(scene)
        assert len(blob) % 4 == 0 and blob[:4] == b"RBT1"
        t = rb.tree.Tree(blob)
        assert t.size == (pytest.approx(scene["width"]), pytest.approx(scene["height"]))
        t.close()
        n += 1
    assert n >= 850


def test_truncated_and_corrupted_streams_are_rejected_not_trusted():
    blob = rb.tree.serialize(_scene("filters__filter__with-mask-on-parent")) if os.path.exists(
        os.path.join(HERE, "golden", "scenes", "filters__filter__with-mask-on-parent.json")) else rb.tree.serialize(
        _scene(sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(HERE, "golden", "scenes", "filters__*.json")))[0]))
    assert _parse(blob) == 0
    assert _parse(b"") != 0 and _parse(b"RBT1") != 0 and _parse(b"XXXX" + blob[4:]) != 0
    assert _parse(blob + b"\0\0\0\0") != 0  # trailing bytes
    rng = np.random.default_rng(1)
    for cut in sorted(set(int(x) for x in rng.integers(8, len(blob) - 4, 200))):
        assert _parse(blob[:cut & ~3]) != 0, cut  # every truncation fails cleanly
    # random word corruption: must never crash; counts / enums / sizes are validated before use
    for _ in range(300):
        b = bytearray(blob)
        i = int(rng.integers(3, len(b) // 4)) * 4
        b[i:i + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes() if rng.random() < 0.5 else b"\xff\xff\xff\x7f"
        _parse(bytes(b))


def test_node_lookup_and_bounding_boxes():
    name = "structure__style__external-CSS" if False else None
    path = os.path.join(HERE, "golden", "extra", "filter-with-transform-on-shape.json")
    with open(path) as f:
        d = json.load(f)
    scene = d["scene"]
    t = rb.tree.Tree(scene)
    g1 = t.node_bbox("g1")
    from tests import svgfront as F
    want = F.find_node(scene["root"], "g1")["abs_layer_bbox"]
    assert g1 == pytest.approx(tuple(want))
    assert t.node_bbox("rect1") is not None and t.node_bbox("frame") is not None
    assert t.node_bbox("") is None and t.node_bbox("nope") is None
    t.close()


def test_rust_shim_writer_covers_the_stream_grammar():
    """shim/resvg-b200/src/lib.rs cannot be compiled here; at least every node / paint / primitive tag of the grammar must
    appear in its writer, in the order the Python writer (exercised above and on the GPU) uses."""
    src = open(os.path.join(os.path.dirname(HERE), "shim", "resvg-b200", "src", "lib.rs")).read()
    assert "0x3154_4252" in src
    for k in ("Blend", "DropShadow", "Flood", "GaussianBlur", "Offset", "Composite", "Merge", "Tile", "Image", "ComponentTransfer",
              "ColorMatrix", "ConvolveMatrix", "Morphology", "DisplacementMap", "Turbulence", "DiffuseLighting", "SpecularLighting"):
        assert f"K::{k}(" in src, k
    import re
    tags = [int(m) for m in re.findall(r"K::\w+\(fe\) => \{\s*w\.u32\((\d+)\)", src)]
    assert tags == sorted(tags) and set(tags) == set(range(17)) - {2} | {2} or len(tags) >= 15
    for prim, tag in rb.tree.PRIM.items():
        assert 0 <= tag <= 16
