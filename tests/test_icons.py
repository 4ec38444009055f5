"""BASELINE config 5 ('icons'): the document generator, the CPU checker's traversal (CPU
tests) and the atlas renderer of resvg_b200/documents.py against the checker, document by document (GPU tests)."""
import numpy as np
import pytest

from resvg_b200 import scenes
from tests import icons_ref


def test_icon_documents_are_seeded_per_document():
    """Document i is a function of ICON_SEED + i alone: any chunking of the batch yields the same documents."""
    a = scenes.icons_docs(0, 40)
    b = scenes.icons_docs(17, 5)
    assert a["n_docs"] == 40 and len(a["doc_first"]) == 41
    n = np.diff(a["doc_first"].astype(np.int64))
    assert n.min() >= 5 and n.max() <= 40
    lo, hi = a["doc_first"][17], a["doc_first"][22]
    assert hi - lo == b["n_paths"]
    pa, pb = a["pt_off"][lo], a["pt_off"][hi]
    assert np.array_equal(a["pts"][pa:pb], b["pts"])
    assert np.array_equal(a["color"][lo:hi], b["color"])
    assert np.array_equal(a["group_opacity"][17:22], b["group_opacity"])
    assert a["pts"].min() > -100 and a["pts"].max() < 356  # centre in the cell, radius <= 96
    big = scenes.icons_docs(0, 4000)
    frac_group = (big["group_first"] >= 0).mean()
    frac_shadow = (big["shadow_sigma"] > 0).mean()
    assert 0.07 < frac_group < 0.13 and 0.03 < frac_shadow < 0.07
    assert not ((big["group_first"] >= 0) & (big["shadow_sigma"] > 0)).any()
    s = big["shadow_sigma"][big["shadow_sigma"] > 0]
    assert s.min() >= 2.0 and s.max() <= 4.0


def test_checker_group_opacity_and_shadow_change_the_result():
    """The checker's traversal: a group with opacity differs from drawing its children directly; a drop shadow adds pixels
    outside the source's alpha."""
    sc = scenes.icons_docs(0, 24)
    paints = icons_ref.prepare(sc)
    full, _ = icons_ref.render_docs(sc, paints)
    flat = dict(sc, group_first=np.full(24, -1), shadow_sigma=np.zeros(24))
    plain, _ = icons_ref.render_docs(flat, paints)
    g = np.nonzero(sc["group_first"] >= 0)[0]
    s = np.nonzero(sc["shadow_sigma"] > 0)[0]
    assert len(g) and len(s)
    for k in range(24):
        same = np.array_equal(full[k], plain[k])
        assert same == (k not in g and k not in s), k
    for k in s:
        assert ((full[k][..., 3] > 0) & (plain[k][..., 3] == 0)).sum() > 50


@pytest.mark.gpu
@pytest.mark.parametrize("first,n,cols,rows", [(0, 70, 8, 9), (1000, 33, 7, 5)])
def test_atlas_renderer_matches_checker_per_document(ctx, first, n, cols, rows):
    from resvg_b200 import documents
    sc = documents.prepare_chunk(first, n)
    atlas = documents.IconAtlas(ctx, cols, rows)
    chunk = atlas.render(sc)
    got = atlas.atlas.download()
    atlas.release(chunk)
    chunk = atlas.prepare(sc)  # the resident form bench.py times: same pixels, any number of times
    for _ in range(2):
        atlas.run(chunk)
    assert np.array_equal(atlas.atlas.download(), got)
    atlas.release(chunk)
    want, _ = icons_ref.render_docs(sc, icons_ref.prepare(sc))
    size = sc["doc_size"]
    kinds = sc["paint_kind"]
    n_exact = 0
    for k in range(n):
        x, y = (k % cols) * size, (k // cols) * size
        cell = got[y:y + size, x:x + size]
        d = np.abs(cell.astype(np.int16) - want[k].astype(np.int16))
        radial = (kinds[sc["doc_first"][k]:sc["doc_first"][k + 1]] == 2).any()
        if not radial:  # lowp only: bit-exact
            assert d.max() == 0, f"doc {first + k}: {int((d > 0).sum())} bytes differ (max {int(d.max())})"
            n_exact += 1
        else:           # two-point conical gradients run in the f32 pipeline: 1/255
            assert d.max() <= 1, f"doc {first + k}: max diff {int(d.max())}"
    assert n_exact >= 2
    # cells beyond the last document stay transparent
    used_rows = (n + cols - 1) // cols
    assert not got[used_rows * size:].any()
    if n % cols:
        assert not got[(used_rows - 1) * size:used_rows * size, (n % cols) * size:].any()


@pytest.mark.gpu
def test_draw_layer_rects_equals_per_rect_draw_layer(ctx):
    import resvg_b200 as rb
    from tests.util import random_premul
    base, src = random_premul(96, 64, 1), random_premul(96, 64, 2, sparse=True)
    rects = np.array([[0, 0, 32, 32], [40, 8, 17, 23], [64, 40, 32, 24], [90, 2, 20, 20], [5, 40, 0, 3]], np.int32)
    op = np.array([0.5, 1.0, 0.25, 0.8, 0.3], np.float32)
    dst = ctx.layer_from(base)
    s = ctx.layer_from(src)
    rb.draw_layer_rects(dst, s, rects, op)  # src_xy = None: same position
    want = ctx.layer_from(base)
    for (x, y, w, h), o in zip(rects, op):
        w, h = min(w, 96 - x), min(h, 64 - y)
        if w <= 0 or h <= 0:
            continue
        sub = ctx.layer_from(np.ascontiguousarray(src[y:y + h, x:x + w]))
        rb.draw_layer(want, sub, int(x), int(y), float(o))
    assert np.array_equal(dst.download(), want.download())


@pytest.mark.gpu
def test_flood_alpha_matches_checker(ctx):
    import resvg_b200 as rb
    from tests.svgfilters import _recolor
    from tests.util import random_premul
    img = random_premul(67, 41, 5)
    for color, opacity in [((0, 0, 0), 0.5), ((255, 128, 3), 1.0), ((12, 200, 99), 0.33), ((255, 255, 255), 0.0)]:
        l = ctx.layer_from(img)
        rb.filters.flood_alpha(color, int(np.ceil(np.float32(opacity) * np.float32(255.0))), l)
        want = img.copy()
        _recolor(icons_ref._Be, want, color, opacity)
        assert np.array_equal(l.download(), want), (color, opacity)


@pytest.mark.gpu
@pytest.mark.parametrize("odd", [False, True])
def test_box_blur_cells_blurs_every_rectangle_as_its_own_pixmap(ctx, oracle, odd):
    import resvg_b200 as rb
    from tests.util import random_premul
    W, H = (301, 150) if odd else (320, 160)
    img = random_premul(W, H, 8)
    if odd:
        rects = [(3, 5, 77, 40), (81, 0, 33, 150), (150, 60, 120, 51), (290, 140, 11, 10)]
    else:
        rects = [(0, 0, 64, 64), (64, 0, 128, 96), (192, 32, 128, 128), (0, 100, 40, 60)]
    sx = [2.0, 3.7, 0.0, 9.0]
    sy = [2.0, 0.0, 4.4, 1.0]   # sigma 0 on an axis: that axis is copied; sigma < 2 still uses box sizes here
    l = ctx.layer_from(img)
    rb.filters.box_blur_cells(rects, sx, sy, l)
    got = l.download()
    want = img.copy()
    for (x, y, w, h), a, b in zip(rects, sx, sy):
        want[y:y + h, x:x + w] = oracle.box_blur(a, b, np.ascontiguousarray(img[y:y + h, x:x + w]))
    assert np.array_equal(got, want)
    # vertical only -> odd number of passes -> the result is copied back cell by cell
    l = ctx.layer_from(img)
    rb.filters.box_blur_cells(rects[:2], [0.0, 0.0], [3.0, 5.0], l)
    want = img.copy()
    for (x, y, w, h), b in zip(rects[:2], [3.0, 5.0]):
        want[y:y + h, x:x + w] = oracle.box_blur(0.0, b, np.ascontiguousarray(img[y:y + h, x:x + w]))
    assert np.array_equal(l.download(), want)


@pytest.mark.gpu
def test_documents_do_not_depend_on_their_chunk_or_cell(ctx):
    """Full-size atlases (32 x 32 cells): documents 1024..1151 rendered as the head of their own chunk and as part of a full
    1024-document chunk at other cells of the atlas give the same pixels (opacity groups and drop shadows included)."""
    from resvg_b200 import documents
    atlas = documents.IconAtlas(ctx)  # 8192 x 8192
    size = 256
    sc_a = documents.prepare_chunk(1024 - 896, 1024)  # documents 128..1151: 1024.. sit at cells 896..1023
    ch = atlas.render(sc_a)
    full = atlas.atlas.download()
    atlas.release(ch)
    sc_b = documents.prepare_chunk(1024, 128)         # the same documents at cells 0..127
    ch = atlas.render(sc_b)
    part = atlas.atlas.download()
    atlas.release(ch)
    assert (sc_b["shadow_sigma"] > 0).any() and (sc_b["group_first"] >= 0).any()
    for k in range(128):
        ca, cb = 896 + k, k
        a = full[(ca // 32) * size:(ca // 32 + 1) * size, (ca % 32) * size:(ca % 32 + 1) * size]
        b = part[(cb // 32) * size:(cb // 32 + 1) * size, (cb % 32) * size:(cb % 32 + 1) * size]
        assert np.array_equal(a, b), f"document {1024 + k}"
    assert not part[4 * size:].any()
