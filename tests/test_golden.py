"""Golden pinning.  tests/golden/scenes/*.json are scene trees produced by tests/svgfront.parse from the reference's
regression corpus (crates/resvg/tests/tests/**.svg), *.png the reference's own golden renders (made by resvg itself).
The script that generated them is tests/golden/make_fixtures.py; tests/golden/CORPUS_RESULTS.md has the whole-corpus
table (861 goldens reproduced).

CPU: the oracle must reproduce every golden at the reference's own criterion (±1 per demultiplied channel, zero
differing pixels — crates/resvg/tests/integration/main.rs:151-226).
GPU: the CUDA path must match the oracle on the same scenes (bit-exact for scenes that stay on the integer pipeline,
±1/255 where f32 stages are involved) and the golden at the reference's criterion.
"""
import glob
import json
import os

import numpy as np
import pytest
from PIL import Image

from tests import svgfront as F

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = sorted(glob.glob(os.path.join(HERE, "golden", "scenes", "*.json")))
IDS = [os.path.basename(s)[:-5] for s in SCENES]


def _load(path):
    with open(path) as f:
        scene = json.load(f)
    gold = np.array(Image.open(path[:-5] + ".png").convert("RGBA"))
    return scene, gold


def test_fixture_set_is_present():
    assert len(SCENES) >= 850


@pytest.mark.parametrize("path", SCENES, ids=IDS)
def test_oracle_reproduces_reference_golden(path):
    from tests.backends import OracleBackend

    scene, gold = _load(path)
    out = F.render_scene(scene, OracleBackend(), 300)
    assert F.diff_pixels(out, gold) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", SCENES, ids=IDS)
def test_gpu_matches_oracle_and_golden(ctx, path):
    from tests.backends import OracleBackend

    scene, gold = _load(path)
    want = F.render_scene(scene, OracleBackend(), 300)   # checker: Python traversal + C oracle
    got = F.render_scene_gpu(scene, ctx, 300)            # product: rb_render (C++ traversal + CUDA)
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    # f32 stages (layer composites with opacity, two-point gradients, lighting powf ...) may differ by one unit
    assert d.max() <= 1, f"max |gpu - oracle| = {d.max()} at {np.argwhere(d > 1)[:3].tolist()}"
    assert F.diff_pixels(got, gold) == 0


# ---- crates/resvg/tests/integration/extra.rs: native size, scale 10, render_node by id -------------------------------------
EXTRA = sorted(glob.glob(os.path.join(HERE, "golden", "extra", "*.json")))
EXTRA_IDS = [os.path.basename(s)[:-5] for s in EXTRA]


def _int_size(w, h):  # Size::to_int_size
    return max(1, int(np.floor(np.float32(w) + np.float32(0.5)))), max(1, int(np.floor(np.float32(h) + np.float32(0.5))))


def extra_target(scene, mode, arg):
    """integration/main.rs:88-100 -> (pixmap w, h, transform) or None when the node has no layer bounding box."""
    if mode == "extra":
        iw, ih = _int_size(scene["width"], scene["height"])
        s = float(arg)
        return int(np.floor(np.float32(iw) * np.float32(s) + np.float32(0.5))), int(np.floor(np.float32(ih) * np.float32(s) + np.float32(0.5))), F.ts_scale(s, s)
    n = F.find_node(scene["root"], arg)
    if n is None:
        return None
    bbox = n.get("abs_layer_bbox") if n["t"] == "g" else n.get("abs_bbox")
    if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
        return None
    w, h = _int_size(bbox[2], bbox[3])
    return w, h, F.IDENT


def render_extra_oracle(scene, mode, arg):
    from tests.backends import OracleBackend
    tgt = extra_target(scene, mode, arg)
    if tgt is None:
        return None
    w, h, ts = tgt
    be = OracleBackend()
    r = F.Renderer(be)
    if mode == "extra":
        return be.to_numpy(r.render(scene, w, h, ts))
    layer = be.new_layer(w, h)
    assert r.render_node_by_id(scene, arg, ts, layer)
    return be.to_numpy(layer)


def _load_extra(path):
    with open(path) as f:
        d = json.load(f)
    return d, np.array(Image.open(path[:-5] + ".png").convert("RGBA"))


def test_extra_fixture_set_is_present():
    assert len(EXTRA) >= 13


@pytest.mark.parametrize("path", EXTRA, ids=EXTRA_IDS)
def test_oracle_reproduces_extra_golden(path):
    d, gold = _load_extra(path)
    out = render_extra_oracle(d["scene"], d["mode"], d["arg"])
    assert out is not None and F.diff_pixels(out, gold) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", EXTRA, ids=EXTRA_IDS)
def test_gpu_extra_matches_oracle_and_golden(ctx, path):
    import resvg_b200 as rb
    d, gold = _load_extra(path)
    scene, mode, arg = d["scene"], d["mode"], d["arg"]
    want = render_extra_oracle(scene, mode, arg)
    w, h, ts = extra_target(scene, mode, arg)
    tree = rb.tree.Tree(scene)
    layer = ctx.layer(w, h)
    if mode == "extra":
        rb.tree.render(tree, ts, layer)
    else:
        bbox = tree.node_bbox(arg)  # node.abs_layer_bounding_box() through the C ABI
        assert bbox is not None and _int_size(bbox[2], bbox[3]) == (w, h)
        assert rb.tree.render_node(tree, arg, ts, layer)
    got = layer.download()
    dd = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert dd.max() <= 1, f"max |gpu - oracle| = {dd.max()}"
    assert F.diff_pixels(got, gold) == 0


@pytest.mark.gpu
def test_render_node_unknown_id_and_render_to_host(ctx):
    """resvg::render_node returns None for an unknown / zero-sized node; resvg_render draws over the caller's pixmap."""
    import resvg_b200 as rb
    d, _ = _load_extra(EXTRA[0])
    scene = d["scene"]
    tree = rb.tree.Tree(scene)
    layer = ctx.layer(16, 16)
    assert rb.tree.render_node(tree, "no-such-node", F.IDENT, layer) is False
    assert tree.node_bbox("no-such-node") is None
    w, h, ts = extra_target(scene, "extra", 1.0)
    host = np.zeros((h, w, 4), np.uint8)
    host[...] = (0, 0, 64, 64)  # premultiplied background kept where nothing is drawn
    rb.tree.render_to_host(ctx, tree, ts, host)
    ref = ctx.layer_from(np.full((h, w, 4), (0, 0, 64, 64), np.uint8))
    rb.tree.render(tree, ts, ref)
    assert np.array_equal(host, ref.download())
