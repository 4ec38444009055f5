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
