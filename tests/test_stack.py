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
