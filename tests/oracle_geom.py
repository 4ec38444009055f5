"""TEMPORARY shim until oracle/stroke.c, dash.c, hairline.c exist: forwards to the product's host geometry."""
import resvg_b200 as rb


def stroke_path(verbs, pts, width, miter_limit, cap, join, res_scale):
    return rb.stroke_path(verbs, pts, width, miter_limit, cap, join, res_scale)


def dash_path(verbs, pts, dash_array, dash_offset, res_scale):
    return rb.dash_path(verbs, pts, dash_array, dash_offset, res_scale)


def hairline_blits(verbs, pts, cap, clip_w, clip_h):
    return rb.hairline_blits(verbs, pts, cap, clip_w, clip_h)
