"""CPU checker for the icon documents of resvg_b200.scenes.icons_docs (BASELINE config 5): the same host traversal as
resvg_b200/documents.py (render.rs:49-143 for these documents), but one document at a time into a pixmap of its own and
with every pixel operation done by the oracle (oracle/raster.c, oracle/filters.c).  Test infrastructure: used by
tests/test_icons_gpu.py and by bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import time

import numpy as np

from tests import oracle_ffi as O
from tests import scenes_loader
from tests import oracle_raster as R
from tests.svgfilters import _recolor

scenes = scenes_loader.load()  # without importing the product package (bench.py --impl reference)


class _Be:  # the slice of the back-end interface _recolor needs
    name = "oracle"

    @staticmethod
    def to_numpy(l):
        return l


def prepare(sc):
    """Oracle paint table for a chunk (outside any timed region)."""
    return scenes.to_paint_array(sc, R.Paint)


def render_docs(sc, paints, docs=None):
    """Renders documents `docs` (indices into the chunk; default all) -> (len(docs), size, size, 4) premultiplied RGBA8 and
    the seconds spent inside the oracle calls."""
    size = sc["doc_size"]
    docs = range(sc["n_docs"]) if docs is None else docs
    out = np.zeros((len(docs), size, size, 4), np.uint8)
    ident = R.ts_arr(R.IDENTITY)
    psz = C.sizeof(R.Paint)
    first, gf, sig = sc["doc_first"], sc["group_first"], sc["shadow_sigma"]

    def fill(px, a, b):
        if b > a:
            R.lib.orc_fill_paths(px.ctypes.data, size, size, b - a, sc["verb_off"].ctypes.data + 4 * a, sc["pt_off"].ctypes.data + 4 * a,
                                 sc["verbs"].ctypes.data, sc["pts"].ctypes.data, C.addressof(paints) + psz * a,
                                 sc["rules"].ctypes.data + a, ident)

    t0 = time.perf_counter()
    for j, k in enumerate(docs):
        px = out[j]
        a, e = int(first[k]), int(first[k + 1])
        if sig[k] > 0:
            src = np.zeros((size, size, 4), np.uint8)
            fill(src, a, e)
            shd = O.box_blur(float(sig[k]), float(sig[k]), src)
            _recolor(_Be, shd, (0, 0, 0), 0.5)  # to_u8() = 128
            shd = O.into_srgb(shd)
            res = np.zeros((size, size, 4), np.uint8)
            R.draw_pixmap(res, 4, 4, shd)
            R.draw_pixmap(res, 0, 0, src)
            R.draw_pixmap(px, 0, 0, res)
        elif gf[k] >= 0:
            g = int(gf[k])
            fill(px, a, g)
            sub = np.zeros((size, size, 4), np.uint8)
            fill(sub, g, e)
            R.draw_pixmap(px, 0, 0, sub, float(sc["group_opacity"][k]))
        else:
            fill(px, a, e)
    return out, time.perf_counter() - t0
