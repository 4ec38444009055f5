"""Host traversal for batches of small documents (BASELINE config 5, SURVEY.md section 8(d) C5 'icons').

This is the part of crates/resvg/src/render.rs that stays on the host, restated for the synthetic icon documents of
``scenes.icons_docs``: per document a list of filled paths, optionally one group with opacity (render.rs:108-133: the
group is rendered into a transparent layer of its own and composited with draw_pixmap at the group's opacity) or one
group with a drop-shadow filter (filter/mod.rs:581-641).  Documents are rendered side by side into one atlas layer:

* pass 1 — ONE batch renders the ungrouped paths of every document into its cell of the atlas (rb_batch_draw_documents);
* pass 2 — ONE batch renders every document's group children into the same cell of a second, transparent atlas;
* the opacity groups of all documents are composited with one launch (rb_draw_layer_rects);
* drop-shadow groups run the reference's primitive sequence for all documents at once on strips of cell-sized
  sub-images: region-wise copies/draws (rb_draw_layer_rects), a blur that is told the cells so that no window sees a
  neighbouring document (rb_filter_box_blur_cells), and pointwise flood / colour-space passes over the whole strip.

Everything is C-ABI calls on device layers; nothing here touches pixels on the host.
"""
import numpy as np

from . import api as rb
from . import _ffi, scenes


class IconAtlas:
    """cols x rows cells of size x size px: the render target for up to cols * rows documents per pass."""

    def __init__(self, ctx, cols=32, rows=32, size=scenes.ICON_SIZE):
        self.ctx, self.cols, self.rows, self.size = ctx, cols, rows, size
        self.atlas = ctx.layer(cols * size, rows * size)
        self.groups = ctx.layer(cols * size, rows * size)
        self._strip, self._strip_rows = None, 0  # source, shadow, result sub-images of the drop-shadow groups

    @property
    def capacity(self):
        return self.cols * self.rows

    def viewports(self, n_docs):
        k = np.arange(n_docs)
        vp = np.empty((n_docs, 4), np.int32)
        vp[:, 0] = (k % self.cols) * self.size
        vp[:, 1] = (k // self.cols) * self.size
        vp[:, 2:] = self.size
        return vp

    def render(self, sc, n_threads=0):
        """Renders the documents of `sc` (scenes.icons_docs; sc["paints"] = ctypes rb_paint array) into the atlas, cell k =
        document k: host edge build + upload + kernels (rb_batch_submit).  Returns the prepared chunk it executed."""
        return self._execute(sc, n_threads, resident=False)

    def prepare(self, sc, n_threads=0):
        """Host build + upload only (rb_batch_prepare): the returned chunk can be executed any number of times with run()."""
        return self._execute(sc, n_threads, resident=True)

    def run(self, chunk):
        """Executes a prepared chunk: kernels only, inputs resident in HBM."""
        self.atlas.fill(0, 0, 0, 0)
        chunk["base"].run()
        if chunk["group"] is not None:
            self.groups.fill(0, 0, 0, 0)
            chunk["group"].run()
            self._composite(chunk)

    def _execute(self, sc, n_threads, resident):
        n = sc["n_docs"]
        assert n <= self.capacity
        vp = self.viewports(n)
        first = sc["doc_first"][:-1].astype(np.int64)
        end = sc["doc_first"][1:].astype(np.int64)
        gf = sc["group_first"]
        shadow = sc["shadow_sigma"] > 0
        # a drop-shadow group holds the whole document; an opacity group holds [group_first, end)
        g_first = np.where(shadow, first, np.where(gf >= 0, gf, end))
        has_group = g_first < end
        chunk = dict(sc=sc, vp=vp, group=None, opacity_docs=np.nonzero(has_group & ~shadow)[0], shadow_docs=np.nonzero(shadow)[0])
        if not resident:
            self.atlas.fill(0, 0, 0, 0)
        b = rb.Batch(self.atlas)
        b.draw_documents(sc, vp, first, g_first - first)
        b.prepare(n_threads) if resident else b.submit(n_threads)
        chunk["base"] = b
        if has_group.any():
            idx = np.nonzero(has_group)[0]
            if not resident:
                self.groups.fill(0, 0, 0, 0)
            g = rb.Batch(self.groups)
            g.draw_documents(sc, vp[idx], g_first[idx], (end - g_first)[idx])
            g.prepare(n_threads) if resident else g.submit(n_threads)
            chunk["group"] = g
            if not resident:
                self._composite(chunk)
        return chunk

    def _composite(self, chunk):
        sc, vp = chunk["sc"], chunk["vp"]
        op = chunk["opacity_docs"]
        if len(op):
            rb.draw_layer_rects(self.atlas, self.groups, vp[op], sc["group_opacity"][op])
        sh = chunk["shadow_docs"]
        if not len(sh):
            return
        # The drop-shadow groups of all documents at once, on strips of cell-sized sub-images (one per shadow document):
        # only the blur looks beyond a pixel, and it is told the cells; the draws are region-wise draw_pixmaps.
        F = rb.filters
        n, size = len(sh), self.size
        strip = self._strips(n)
        src, shd, out = strip
        cells = np.zeros((n, 4), np.int32)
        cells[:, 0] = (np.arange(n) % self.cols) * size
        cells[:, 1] = (np.arange(n) // self.cols) * size
        cells[:, 2:] = size
        for l in strip:
            l.fill(0, 0, 0, 0)
        rb.draw_layer_rects(src, self.groups, cells, 1.0, src_xy=vp[sh, :2])  # the group's own layer (render.rs:108)
        shd.copy_from(src)                                                    # filter/mod.rs:594
        F.box_blur_cells(cells, sc["shadow_sigma"][sh], sc["shadow_sigma"][sh], shd)
        F.flood_alpha((0, 0, 0), 128, shd)                                    # :606-617
        F.into_srgb(shd)                                                      # :619-622 (color-interpolation-filters = sRGB)
        off = cells.copy()
        off[:, :2] += 4
        off[:, 2:] -= 4
        rb.draw_layer_rects(out, shd, off, 1.0, src_xy=cells[:, :2])          # :624-631 draw_pixmap(dx, dy), clipped to the image
        rb.draw_layer_rects(out, src, cells, 1.0)                             # :633-640
        dst = np.concatenate([vp[sh, :2], cells[:, 2:]], axis=1)
        rb.draw_layer_rects(self.atlas, out, dst, 1.0, src_xy=cells[:, :2])   # render.rs:133

    def _strips(self, n):
        rows = (n + self.cols - 1) // self.cols
        if self._strip_rows < rows:
            self._strip = [self.ctx.layer(self.cols * self.size, rows * self.size) for _ in range(3)]
            self._strip_rows = rows
        return self._strip

    @staticmethod
    def release(chunk):
        for k in ("base", "group"):
            if chunk.get(k) is not None:
                chunk[k].close()
                chunk[k] = None


def prepare_chunk(first_doc, n_docs):
    """icons_docs + the ctypes paint table the batch API consumes."""
    sc = scenes.icons_docs(first_doc, n_docs)
    sc["paints"] = scenes.to_paint_array(sc, _ffi.Paint)
    return sc
