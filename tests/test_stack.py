"""BASELINE config 4 ('stack4k', SURVEY.md section 8(d) C4: nested groups with opacity, luminance masks, clip-paths incl.
nested ones, and pattern fills) as a parity case: the SVG of resvg_b200.scenes.stack_svg goes through the test-side front
end (tests/svgfront.py = usvg + render.rs/clip.rs/mask.rs/path.rs traversal) once per back end."""
import numpy as np
import pytest

from resvg_b200 import scenes
from tests import svgfront as F


def _scene(size, levels, inset):
    return F.parse(scenes.stack_svg(size, levels, inset=inset))


def test_stack_scene_structure_and_checker_render():
    sc = _scene(256, 9, 6.0)
    depth, g, kinds = 0, sc["root"], []
    while True:
        subs = [c for c in g["children"] if c["t"] == "g"]
        if not subs:
            break
        g = subs[-1]
        if not g.get("isolate"):  # the root wrapper of the <svg> element
            continue
        kinds.append(("mask" if g.get("mask") else "") + ("clip" if g.get("clip") else ""))
        depth += 1
    assert depth == 9
    assert kinds[0] == "mask" and kinds[1] == "clip" and kinds[4] == "mask" and kinds[5] == "clip" and kinds[3] == ""
    from tests.backends import OracleBackend
    out = F.render_scene(sc, OracleBackend(), 256)
    assert out.shape == (256, 256, 4) and out[..., 3].max() > 100
    # the outermost level is masked and faded: nothing is fully opaque
    assert out[..., 3].max() < 255


@pytest.mark.gpu
@pytest.mark.parametrize("size,levels,inset", [(512, 16, 8.0), (300, 8, 5.5)])
def test_gpu_stack_matches_checker(ctx, size, levels, inset):
    from tests.backends import OracleBackend
    sc = _scene(size, levels, inset)
    want = F.render_scene(sc, OracleBackend(), size)
    got = F.render_scene_gpu(sc, ctx, size, via="submit")
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    # layer composites with opacity and luminance masks run in f32: one unit per stage
    assert d.max() <= 1, f"max |gpu - checker| = {d.max()}, {int((d > 1).sum())} bytes"
    assert (d > 0).mean() < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("size,levels,inset,cuts", [(512, 16, 8.0, (0, 200, 328, 512)), (300, 8, 5.5, (0, 8, 96, 300))])
def test_gpu_tree_rendered_in_canvas_strips_equals_the_whole_render(ctx, size, levels, inset, cuts):
    """one document cut into canvas strips (rb_render_strip, one strip per GPU) — nested groups with opacity,
    luminance masks, clip-paths and patterns crossing the cuts — holds exactly the pixels of the whole-canvas rb_render."""
    import resvg_b200 as rb

    sc = _scene(size, levels, inset)
    pw, ph, ts = F.target_for(sc, size)
    tree = rb.tree.Tree(sc)
    whole = ctx.layer(pw, ph)
    rb.tree.render(tree, ts, whole)
    want = whole.download()
    got = np.zeros_like(want)
    rows = [min(c, ph) for c in cuts]
    for y0, y1 in zip(rows[:-1], rows[1:]):
        strip = ctx.layer(pw, y1 - y0)
        rb.tree.render_strip(tree, ts, pw, ph, y0, strip)
        got[y0:y1] = strip.download()
    tree.close()
    assert rows[-1] == ph
    assert np.array_equal(got, want), f"{int((got != want).any(axis=-1).sum())} pixels differ between the strips and the whole render"


@pytest.mark.gpu
def test_gpu_corpus_scenes_in_strips(ctx):
    """The same on fixtures of the regression corpus that carry filters, masks, clip-paths and nested svg images."""
    import glob
    import json
    import os

    import resvg_b200 as rb

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenes")
    names = sorted(glob.glob(os.path.join(root, "*.json")))
    picked = [n for n in names if any(k in os.path.basename(n) for k in ("feGaussianBlur", "mask__", "clipPath__", "pattern__", "image__", "opacity"))][::7][:40]
    assert len(picked) >= 20
    checked = 0
    for n in picked:
        with open(n) as f:
            sc = json.load(f)
        pw, ph, ts = F.target_for(sc, 300)
        if ph < 24:
            continue
        tree = rb.tree.Tree(sc)
        whole = ctx.layer(pw, ph)
        rb.tree.render(tree, ts, whole)
        want = whole.download()
        got = np.zeros_like(want)
        cut = (ph // 3) | 1
        for y0, y1 in ((0, cut), (cut, ph)):
            strip = ctx.layer(pw, y1 - y0)
            rb.tree.render_strip(tree, ts, pw, ph, y0, strip)
            got[y0:y1] = strip.download()
        tree.close()
        assert np.array_equal(got, want), f"{os.path.basename(n)}: {int((got != want).any(axis=-1).sum())} pixels differ"
        checked += 1
    assert checked >= 20
