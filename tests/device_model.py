"""Pure-numpy model of the device coverage algorithm in resvg_b200/csrc/raster.cu (k_raster_tiles), fed by the
product's own host edge builder through rb_debug_build_edges.  It lets the CPU suite check the host geometry and
the data-parallel coverage formulation against the sequential oracle without a GPU.  Small canvases only."""
import ctypes as C
import functools

import numpy as np

from resvg_b200 import _ffi

IDENTITY = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


def build_edges(verbs, pts, aa, w, h, ts=IDENTITY, cap=1 << 18):
    v = np.ascontiguousarray(verbs, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros((cap, 5), np.int32)
    meta = np.zeros((cap, 2), np.int32)
    geom = np.zeros(7, np.int32)
    n = _ffi.lib.rb_debug_build_edges(v.ctypes.data, len(v), p.ctypes.data, len(p), 1 if aa else 0, w, h,
                                      (C.c_float * 6)(*ts), out.ctypes.data, meta.ctypes.data, cap, geom.ctypes.data)
    assert n >= 0, n
    return out[:n].copy(), meta[:n].copy(), geom


def _i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _x_at(e, y):
    return _i32(int(e[0]) + (y - int(e[2])) * int(e[1]))


def _round16(x):
    return _i32(x + 0x8000) >> 16


def _walker_less(edges, meta, ia, ib, y):
    """Mirrors walker_less() in raster.cu."""
    A, B = edges[ia], edges[ib]
    for _ in range(2):
        xa, xb = _x_at(A, y), _x_at(B, y)
        if xa != xb:
            return xa < xb
        fya, fyb = int(A[2]), int(B[2])

        def state(e, i, x, fy):
            if y > fy:
                return 0, _i32(x - int(e[1]))
            if meta[i][0] >= 0:
                pe = edges[meta[i][0]]
                return 0, _x_at(pe, int(pe[3]))
            return (2 if meta[i][1] else 1), 0

        ka, pa = state(A, ia, xa, fya)
        kb, pb = state(B, ib, xb, fyb)
        if ka == 0 and kb == 0:
            if pa != pb:
                return pa < pb
            if y > fya and y > fyb:
                y = max(fya, fyb)
                continue
            return ia < ib
        if ka == 0:
            return kb == 1
        if kb == 0:
            return ka == 2
        return ia < ib
    return ia < ib


def _exact_break(edges, meta, y, target, w_before):
    cs = []
    for i, e in enumerate(edges):
        if int(e[2]) <= y <= int(e[3]) and _round16(_x_at(e, y)) == target:
            cs.append(i)
    cs.sort(key=functools.cmp_to_key(
        lambda a, b: -1 if _walker_less(edges, meta, a, b, y) else (1 if _walker_less(edges, meta, b, a, y) else 0)))
    w = w_before
    for i in cs:
        w += int(edges[i][4])
        if w == 0:
            return True
    return False


def coverage(verbs, pts, w, h, rule="nonzero", aa=True, ts=IDENTITY):
    edges, meta, geom = build_edges(verbs, pts, aa, w, h, ts)
    cov = np.zeros((h, w), np.uint8)
    if len(edges) == 0:
        return cov
    sx, sy, sw, sh, shift = [int(v) for v in geom[:5]]
    S = 1 << shift
    lo, hi = sx * S, (sx + sw) * S
    for prow in range(sy, sy + sh):
        acc = np.zeros(w, np.int64)
        for sub in range(S):
            y = prow * S + sub
            pos = np.zeros(w * S + 1, np.int64)
            neg = np.zeros(w * S + 1, np.int64)
            for e in edges:
                if int(e[2]) <= y <= int(e[3]):
                    p = max(_round16(_x_at(e, y)), lo)
                    if p < hi:
                        (pos if e[4] > 0 else neg)[p] += 1
            W = np.cumsum(pos - neg)[: w * S]
            inside = (W & 1) != 0 if rule == "evenodd" else W != 0
            for px in range(sx, sx + sw):
                bits = inside[px * S:(px + 1) * S]
                if shift == 0:
                    acc[px] += 255 if bits[0] else 0
                    continue
                add = 16 * int(bits.sum())
                if sub == 3 and bits.all():
                    brk = False
                    for k in range(1, 4):
                        q = px * 4 + k
                        if pos[q] == 0 and neg[q] == 0:
                            continue
                        if rule == "evenodd":
                            brk = True
                        elif pos[q] and neg[q]:
                            brk = brk or _exact_break(edges, meta, y, q, int(W[q - 1]))
                        elif (int(W[q - 1]) < 0) != (int(W[q]) < 0):
                            brk = True
                    add = 64 if brk else 63
                acc[px] += add
        cov[prow] = np.minimum(acc, 255).astype(np.uint8)
    return cov
