#!/usr/bin/env python
"""bench.py — Mpixels/s rendered on the BASELINE.json workloads (resvg pixel hot path on B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload paths8k|icons|filters8k|stack4k]
                    [--shard documents|strips]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload `paths8k` (BASELINE.json configs[1]): one 8192x8192 canvas, 100k random
cubic/quad/line paths, non-zero + even-odd, fills + strokes (dashes, hairlines), solid + linear + radial paints.  A *step*
is one full render of the scene (clear + every path).  With N GPUs every rank renders its own scene (document-parallel,
weak scaling, no collective on the data path); `value` = N * canvas Mpx / max-over-ranks device time.  `--shard strips`
cuts ONE scene into N canvas strips instead (strong scaling; the strips are bit-identical to the whole-canvas render).

The other configurations print the same JSON line: `icons` (configs[4]: 100 000 documents of 256x256 rendered into
atlases, sharded by atlas chunk), `filters8k` (configs[2]: the 10-primitive filter chain over an 8192x8192 layer) and
`stack4k` (configs[3]: 64 nested groups with masks / clip-paths / patterns: one rb_render call per document, traversal in
C++ inside the library).  The default run (`paths8k`, N = 1) also carries them as sub-records under `configs`, and a
`parity` block per record: the GPU result against the CPU checker's on the sample the cpu_baseline leg renders anyway.

Nothing under tests/ or oracle/ is imported outside the cpu_baseline / parity legs and `--impl reference`; the reference
arm imports nothing of resvg_b200 (its only native library is oracle/liboracle.so).

Printed JSON keys follow the driver's contract; see DESIGN.md §Measurement for how each number is obtained.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, n_paths, seed)
    "paths8k": (8192, 8192, 100_000, 0x5EED0002),
    "paths2k": (2048, 2048, 6_000, 0x5EED0002),  # smoke-sized variant for quick checks (not a bench line)
    # BASELINE configs[4] / C5: 100 000 documents of 256x256 px, 1024 per 8192x8192 atlas, sharded by atlas chunk over the GPUs
    "icons": (8192, 8192, 100_000, 0x5EED0005),
    # BASELINE configs[2] / C3 (ii): the filter chain over one 8192x8192 layer holding 1000 C2-style shapes
    "filters8k": (8192, 8192, 1_000, 0x5EED0003),
    # BASELINE configs[3] / C4: 64 nested groups (opacity, luminance masks, clip-paths, patterns) on a 4096x4096 canvas
    "stack4k": (4096, 4096, 64, 0x5EED0004),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_lib():
    from tests import oracle_raster as R  # the CPU checker: only cpu_baseline / --impl reference use it
    return R


def parity_record(got, want, what):
    """GPU result against the CPU checker's on the same input: the largest channel difference and how many pixels differ."""
    d = np.abs(np.asarray(got).astype(np.int16) - np.asarray(want).astype(np.int16))
    per_px = d.reshape(-1, 4).max(axis=1)
    return {"max_abs": int(d.max()) if d.size else 0, "differing_px": int((per_px > 0).sum()), "px_over_1": int((per_px > 1).sum()),
            "compared_px": int(per_px.size), "what": what}


def prepare_sample(R, scene, n_sample):
    """Host geometry of the first n_sample draws for the CPU arm: dash / stroke / hairline-walk every stroke draw with the
    ORACLE's own stroker, dasher and hairline walker (oracle/stroke.c, dash.c, hairline.c through tests/geom.py — nothing of
    libresvg_b200.so), and pack the outlines for the oracle's bulk fill.  Returns the prepared arrays and t_geom, the seconds
    spent INSIDE the C geometry calls (the Python loop around them is not CPU-renderer work)."""
    from tests import geom as rb          # oracle geometry under the names used below
    from tests import scenes_loader
    scenes = scenes_loader.load()
    sub = scenes.subset(scene, n_sample)
    n = sub["n_paths"]
    w, h = scene["width"], scene["height"]
    t_geom = 0.0
    sw = sub["stroke_width"]
    hair = {}  # draw index -> (blits, modulated alpha)
    if (sw > 0).any():
        verbs, pts, voff, poff = [], [], [0], [0]
        caps, joins = ["butt", "round", "square"], ["miter", "miter-clip", "round", "bevel"]
        for i in range(n):
            v = sub["verbs"][sub["verb_off"][i]:sub["verb_off"][i + 1]]
            p = sub["pts"][sub["pt_off"][i]:sub["pt_off"][i + 1]]
            if sw[i] > 0:
                t0 = time.perf_counter()
                src = (v, p)
                if "n_dash" in sub and sub["n_dash"][i] > 0:
                    src = rb.dash_path(v, p, sub["dash"][i][: sub["n_dash"][i]], 0.0, 1.0)
                out = None
                if src is not None and sub["anti_alias"][i] and sw[i] <= 1.0:
                    # treat_as_hairline (identity transform: both mapped width vectors have length w)
                    parts = []
                    for ty in range(0, h, 8191):  # DrawTiler
                        for tx in range(0, w, 8191):
                            bl = rb.hairline_blits(src[0], src[1] - np.float32([tx, ty]), caps[sub["stroke_cap"][i]],
                                                   min(w - tx, 8191), min(h - ty, 8191))
                            if len(bl):
                                bl[:, 0] += tx
                                bl[:, 1] += ty
                                parts.append(bl)
                    scale = int(np.float32(sw[i]) * np.float32(256.0))
                    hair[i] = (np.concatenate(parts) if parts else np.zeros((0, 3), np.int32),
                               np.float32((255 * scale) >> 8) / np.float32(255.0) if sw[i] != 1.0 else np.float32(1.0))
                elif src is not None:
                    out = rb.stroke_outline(src[0], src[1], float(sw[i]), float(sub["stroke_miter"][i]),
                                            caps[sub["stroke_cap"][i]], joins[sub["stroke_join"][i]], 1.0)
                t_geom += time.perf_counter() - t0
                if out is None:
                    v, p = v[:0], p[:0]
                else:
                    v, p = out
                sub["rules"][i] = 0
            verbs.append(v); pts.append(p)
            voff.append(voff[-1] + len(v)); poff.append(poff[-1] + len(p))
        sub["verbs"] = np.ascontiguousarray(np.concatenate(verbs), np.uint8)
        sub["pts"] = np.ascontiguousarray(np.concatenate(pts), np.float32)
        sub["verb_off"] = np.array(voff, np.uint32)
        sub["pt_off"] = np.array(poff, np.uint32)
    paints = scenes.to_paint_array(sub, R.Paint)
    for i, (blits, opacity) in hair.items():  # stroke draws are solid: Shader::apply_opacity scales the colour's alpha
        paints[i].color[3] = float(min(max(np.float32(paints[i].color[3]) * opacity, np.float32(0)), np.float32(1)))
    return dict(sub=sub, paints=paints, hair=hair, n=n, w=w, h=h, t_geom=t_geom)


def render_prepared(R, prep, px):
    """The oracle's pixel work over a prepared sample, in painter's order: bulk fills between the hairline draws, whose
    blits are blended one by one.  Returns the seconds it took (pure C apart from one Python iteration per hairline)."""
    sub, paints, hair, n, w, h = prep["sub"], prep["paints"], prep["hair"], prep["n"], prep["w"], prep["h"]
    psz = C.sizeof(R.Paint)
    ident = R.ts_arr(R.IDENTITY)

    def fill_run(a, b):
        if b > a:
            R.lib.orc_fill_paths(px.ctypes.data, w, h, b - a, sub["verb_off"].ctypes.data + 4 * a, sub["pt_off"].ctypes.data + 4 * a,
                                 sub["verbs"].ctypes.data, sub["pts"].ctypes.data, C.addressof(paints) + psz * a,
                                 sub["rules"].ctypes.data + a, ident)

    t0 = time.perf_counter()
    start = 0
    for i in sorted(hair):
        fill_run(start, i)
        blits = hair[i][0]
        if len(blits):
            R.lib.orc_blit_coverage(px.ctypes.data, w, h, len(blits), blits.ctypes.data, C.byref(paints[i]), ident)
        start = i + 1
    fill_run(start, n)
    return time.perf_counter() - t0


def cpu_render_sample(R, scene, n_sample, canvas=None):
    """Oracle (CPU restatement of the reference path) over the first n_sample draws; returns seconds of CPU work: host
    dashing / stroking / hairline walking of the stroke draws + the oracle's fill of every outline and its blend of
    every hairline blit, in painter's order."""
    prep = prepare_sample(R, scene, n_sample)
    px = canvas if canvas is not None else np.zeros((prep["h"], prep["w"], 4), np.uint8)
    return render_prepared(R, prep, px) + prep["t_geom"]


def kernel_table(rb, ctx, layer, W, H, peak, reps=3):
    """Device time of every other kernel of the path on a W x H layer holding the rendered scene (CUDA events on the
    library's stream), with its algorithmic bytes per pixel -> GB/s and fraction of the HBM peak.
    Inputs are 256 MiB layers (> L2), so no flush is needed between repetitions."""
    F = rb.filters
    a, b = ctx.layer(W, H), ctx.layer(W, H)
    a.copy_from(layer)
    b.copy_from(layer)
    mask = rb.Mask.from_layer(layer, "alpha")
    light = rb.make_light("distant", azimuth=30.0, elevation=40.0)
    ident = [rb.make_transfer("gamma", amplitude=1.0, exponent=0.9, offset=0.01)] * 3 + [rb.make_transfer("identity")]
    rows = [
        ("k_fill_u32 (Pixmap::fill)", 4, lambda: a.fill(0, 0, 0, 0)),
        ("layer copy (Pixmap::clone)", 8, lambda: a.copy_from(layer)),
        ("k_pointwise<demultiply>", 8, lambda: F.demultiply_alpha(a)),
        ("k_pointwise<multiply>", 8, lambda: F.multiply_alpha(a)),
        ("k_cs_convert (into_linear_rgb)", 8, lambda: F.into_linear_rgb(a)),
        ("k_pointwise<color_matrix saturate>", 8, lambda: F.color_matrix("saturate", [0.5], a)),
        ("k_lut4 (component_transfer gamma)", 8, lambda: F.component_transfer(ident, a)),
        ("box_blur sigma 4 (5 V + 5 H passes)", 80, lambda: F.box_blur(4.0, 4.0, a)),
        ("box_blur sigma 20 (5 V + 5 H passes)", 80, lambda: F.box_blur(20.0, 20.0, a)),
        ("box_blur sigma 20, horizontal only (5 passes)", 40, lambda: F.box_blur(20.0, 0.0, a)),
        ("box_blur sigma 20, vertical only (5 passes)", 40, lambda: F.box_blur(0.0, 20.0, a)),
        ("iir_blur sigma 1.5, bit-exact f64 (8 B/px minimum; the 16 sequential f64 plane sweeps move ~1 KB/px)", 8, lambda: F.iir_blur(1.5, 1.5, a)),
        ("iir_blur_fast sigma 1.5, f32 segments + halos, <= 1/255 (8 B/px minimum; 40 B/px with the f32 intermediate)", 8, lambda: F.iir_blur_fast(1.5, 1.5, a)),
        ("morphology dilate r=3 (H + V)", 16, lambda: F.morphology("dilate", 3.0, 3.0, a)),
        ("k_convolve 3x3", 8, lambda: F.convolve_matrix([0, -1, 0, -1, 5, -1, 0, -1, 0], 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, a)),
        ("k_arithmetic", 12, lambda: F.arithmetic(0.1, 0.5, 0.5, 0.0, layer, b, a)),
        ("k_displace", 12, lambda: F.displacement_map(0, 1, 12.0, 1.0, 1.0, layer, b, a)),
        ("k_lighting diffuse distant", 8, lambda: F.diffuse_lighting(2.0, 1.0, (255, 255, 255), light, layer, a)),
        ("k_turbulence 2 octaves", 4, lambda: F.turbulence(0.0, 0.0, 1.0, 1.0, 0.01, 0.01, 2, 1, False, True, a)),
        ("k_draw_layer source_over", 12, lambda: rb.draw_layer(a, layer, 0, 0, 0.75, "source_over")),
        ("k_mask_from_layer luminance", 5, lambda: rb.Mask.from_layer(layer, "luminance")),
        ("k_apply_mask", 9, lambda: rb.apply_mask(a, mask)),
    ]
    def c3_chain():
        # blur 8 -> dilate 3 -> sharpen 3x3 -> arithmetic with turbulence(0.02, 3 oct) -> diffuse
        # lighting (distant 45/60) -> blur 64, in linearRGB
        F.into_linear_rgb(a)
        F.box_blur(8.0, 8.0, a)
        F.morphology("dilate", 3.0, 3.0, a)
        F.convolve_matrix([0, -1, 0, -1, 5, -1, 0, -1, 0], 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, a)
        F.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, 3, 7, False, False, b)
        F.multiply_alpha(b)
        F.arithmetic(0.5, 0.5, 0.5, 0.0, a, b, c)
        F.diffuse_lighting(5.0, 1.0, (255, 255, 255), rb.make_light("distant", azimuth=45.0, elevation=60.0), c, a)
        F.box_blur(64.0, 64.0, a)
        F.into_srgb(a)

    c = ctx.layer(W, H)
    rows.append(("C3 filter chain (blur 8, dilate 3, sharpen, arithmetic x turbulence, diffuse light, blur 64; 10 primitives)",
                 8 + 80 + 16 + 8 + 4 + 8 + 12 + 8 + 80 + 8, c3_chain))
    out = []
    for name, bpp, fn in rows:
        fn()
        ctx.timer_begin()
        for _ in range(reps):
            fn()
        ms = ctx.timer_end() / reps
        gbs = bpp * W * H / (ms * 1e-3) / 1e9
        rec = {"kernel": name, "ms": round(ms, 4), "bytes_per_px": bpp, "GB/s": round(gbs, 1), "frac": round(gbs / peak, 4)}
        if name.startswith("box_blur") and "only" not in name:
            # report the box blur against the reference's 10-pass structure (80 B/px) AND against the
            # 16 B/px minimum of two fused-axis passes (not reachable bit-exactly: every pass quantises to u8)
            rec["frac_of_16B_minimum"] = round(16 * W * H / (ms * 1e-3) / 1e9 / peak, 4)
        out.append(rec)
    return out


ICON_CHUNK = 1024  # documents per atlas (32 x 32 cells of 256 px)


def icon_chunks_for_rank(n_docs, rank, world):
    """[(first_doc, n)] of the atlas chunks this rank renders: chunk c -> rank c mod world (document-parallel, no collective)."""
    chunks = [(f, min(ICON_CHUNK, n_docs - f)) for f in range(0, n_docs, ICON_CHUNK)]
    return chunks[rank::world]


def run_icons(args, rank, local_rank, world, torch, dist):
    """Workload `icons` (BASELINE configs[4]): a step = every one of the 100 000 documents rendered once.
    value: all chunks' batches resident in HBM, kernels only, CUDA events.  e2e: record + host edge build + H2D + kernels +
    D2H of every atlas into pinned memory through the C ABI, downloads overlapped with the next atlas (two atlases)."""
    import resvg_b200 as rb
    from resvg_b200 import documents, shard
    W, H, n_docs, _ = WORKLOADS["icons"]
    n_docs = int(args.docs) if args.docs else n_docs
    ctx = rb.Context(local_rank)
    n_threads = 0 if world == 1 else shard.host_threads(world)
    mine = icon_chunks_for_rank(n_docs, rank, world)
    t0 = time.perf_counter()
    scs = [documents.prepare_chunk(f, n) for f, n in mine]  # synthetic documents: generated outside every timed region
    gen_s = time.perf_counter() - t0
    my_docs = sum(n for _, n in mine)
    atl = [documents.IconAtlas(ctx), documents.IconAtlas(ctx)]
    chunks = [atl[0].prepare(sc, n_threads) for sc in scs]
    ctx.synchronize()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def step():
        for ch in chunks:
            atl[0].run(ch)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms_step = ctx.timer_end() / args.steps
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0

    # dominant kernel (k_raster_warp) alone, and its algorithmic bytes from the kernel's own pixel counters
    ms_kernel = ms_pre = 0.0
    alg_bytes = 0
    stats = dict(draws=0, edges=0, pairs=0, upload_bytes=0)
    for ch in chunks:
        for b in (ch["base"], ch["group"]):
            if b is None:
                continue
            b.run()
            pre, ras = ctx.last_run_ms()
            ms_pre += pre
            ms_kernel += ras
            n_rmw, n_store = b.run_counting()
            st = b.stats()
            alg_bytes += 8 * n_rmw + 4 * n_store + 16 * st["edges"]
            for k in stats:
                stats[k] += st[k]
    for ch in chunks:
        atl[0].release(ch)
    ctx.synchronize()

    # end to end
    pinned = [rb.PinnedBuffer(W * H * 4), rb.PinnedBuffer(W * H * 4)]
    e2e_h2d = 0

    def e2e_step():
        nonlocal e2e_h2d
        prev = [None, None]
        h2d = 0
        for c, sc in enumerate(scs):
            a = atl[c & 1]
            a.atlas.download_end()  # the download that last read this atlas
            if prev[c & 1] is not None:
                a.release(prev[c & 1])
            ch = a.render(sc, n_threads)
            a.atlas.download_begin(pinned[c & 1].array.ctypes.data)
            h2d += ch["base"].stats()["upload_bytes"] + (ch["group"].stats()["upload_bytes"] if ch["group"] is not None else 0)
            prev[c & 1] = ch
        for k in (0, 1):
            atl[k].atlas.download_end()
            if prev[k] is not None:
                atl[k].release(prev[k])
        e2e_h2d = h2d

    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    ms_step, ms_kernel, e2e_s = shard.max_over_ranks([ms_step, ms_kernel, e2e_s], world, f"cuda:{local_rank}")
    if rank != 0:
        return
    total_mpx = n_docs * 256 * 256 / 1e6
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    out = {
        "metric": "Mpixels/s rendered", "value": total_mpx / (ms_step * 1e-3), "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16 fixed point + f32", "data": "synthetic",
        "config": {"workload": "icons", "documents": n_docs, "document_px": [256, 256], "atlas": [W, H], "documents_per_atlas": ICON_CHUNK,
                   "mix": "5-40 filled paths per document, radius log-U[4,96], 80% solid / 10% linear / 10% radial, 95% AA; "
                          "10% of the documents with one opacity group, 5% with a drop shadow (sigma U[2,4])",
                   "sharding": "atlas chunk c (1024 documents) -> GPU c mod N, no collective",
                   "l2": "every atlas pass touches 2 x 256 MiB layers (> 126 MB L2); rank 0 holds %d chunks" % len(mine),
                   "rank0": dict(stats, documents=my_docs, generate_s=round(gen_s, 2))},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_raster_warp", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                     "kernel_ms": ms_kernel, "prepass_ms": ms_pre,
                     "model": "8 B per blended px + 4 B per opaque-stored px + 16 B per line edge, summed over rank 0's batches"},
        "e2e": {"value": total_mpx / e2e_s, "unit": "Mpx/s", "h2d_bytes_per_step": int(e2e_h2d),
                "d2h_bytes_per_step": len(mine) * W * H * 4, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "note": "per atlas: record + host edge build + H2D + kernels, D2H of the atlas overlapped with the next atlas"},
    }
    if world == 1 and not args.no_cpu_baseline:
        from tests import icons_ref  # the CPU checker: cpu_baseline only
        n_s = max(16, min(n_docs, int(args.cpu_sample) // 8))
        sc = scs[0] if scs[0]["n_docs"] >= n_s else scs[0]
        n_s = min(n_s, sc["n_docs"])
        ref_docs, dt = icons_ref.render_docs(sc, icons_ref.prepare(sc), range(n_s))
        out["cpu_baseline"] = {"value": n_s * 0.065536 / dt, "unit": "Mpx/s", "cores": 1, "kind": "port",
                               "sample": f"documents 0..{n_s - 1} of the same batch, one at a time, {dt:.2f} s; oracle restatement of "
                                         "the resvg/tiny-skia CPU path"}
        # parity: the same documents out of the GPU atlas of chunk 0, cell by cell, against the checker's pixmaps
        a = atl[0]
        ch = a.render(scs[0], n_threads)
        got = a.atlas.download()
        a.release(ch)
        n_p = min(n_s, 128)
        out["parity"] = parity_record(np.stack([got[(k // a.cols) * 256:(k // a.cols) * 256 + 256, (k % a.cols) * 256:(k % a.cols) * 256 + 256]
                                                for k in range(n_p)]), ref_docs[:n_p],
                                      f"documents 0..{n_p - 1}: GPU atlas cells vs the CPU checker's per-document pixmaps")
    return out


def run_filters8k(args, rank, local_rank, world, torch, dist):
    """Workload `filters8k` (BASELINE configs[2]): a step = the 10-primitive filter chain over one
    8192x8192 layer (every rank its own layer).  value: layer resident; e2e: upload of the source layer, chain, download."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi, scenes, shard
    W, H, n_paths, seed = WORKLOADS["filters8k"]
    ctx = rb.Context(local_rank)
    scene = scenes.paths_scene(W, H, n_paths, shard.scene_seed(seed, rank))
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    src = ctx.layer(W, H)
    b = rb.Batch(src)
    b.fill_paths(scene)
    b.submit(0 if world == 1 else shard.host_threads(world))
    b.close()
    F = rb.filters
    a, t, c = ctx.layer(W, H), ctx.layer(W, H), ctx.layer(W, H)
    light = rb.make_light("distant", azimuth=45.0, elevation=60.0)
    sharpen = [0, -1, 0, -1, 5, -1, 0, -1, 0]

    def chain():  # reads a, t, c at call time: the parity leg re-points them at crop-sized layers
        F.into_linear_rgb(a)
        F.box_blur(8.0, 8.0, a)
        F.morphology("dilate", 3.0, 3.0, a)
        F.convolve_matrix(sharpen, 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, a)
        F.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, 3, 7, False, False, t)
        F.multiply_alpha(t)
        F.arithmetic(0.5, 0.5, 0.5, 0.0, a, t, c)
        F.diffuse_lighting(5.0, 1.0, (255, 255, 255), light, c, a)
        F.box_blur(64.0, 64.0, a)
        F.into_srgb(a)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def step():
        a.copy_from(src)
        chain()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms_step = ctx.timer_end() / args.steps
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0
    # the chain's dominant primitive alone: the two box blurs (20 of the chain's passes)
    ctx.timer_begin()
    for _ in range(args.steps):
        F.box_blur(8.0, 8.0, a)
        F.box_blur(64.0, 64.0, a)
    ms_blur = ctx.timer_end() / args.steps

    host_src = src.download()
    pin_in, pin_out = rb.PinnedBuffer(W * H * 4), rb.PinnedBuffer(W * H * 4)
    pin_in.array[:] = host_src.reshape(-1)

    def e2e_step():
        a.upload_ptr(pin_in.array.ctypes.data)
        chain()
        a.download_ptr(pin_out.array.ctypes.data)

    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    ms_step, ms_blur, e2e_s = shard.max_over_ranks([ms_step, ms_blur, e2e_s], world, f"cuda:{local_rank}")
    if rank != 0:
        return
    mpx = W * H / 1e6
    peak, peak_src = measured_peaks()
    blur_bytes = 160 * W * H  # 2 blurs x 10 passes x 8 B/px (the reference's pass structure)
    out = {
        "metric": "Mpixels/s rendered", "value": shard.aggregate_throughput(mpx, world, ms_step * 1e-3), "unit": "Mpx/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 + f32/f64", "data": "synthetic",
        "config": {"workload": "filters8k", "canvas": [W, H], "source": "1000 C2-style shapes (seed 0x5EED0003)",
                   "chain": "linearRGB: blur 8 -> dilate 3 -> sharpen 3x3 -> arithmetic(0.5,0.5,0.5,0) with turbulence(0.02, 3 oct) -> "
                            "diffuse lighting (distant 45/60, surfaceScale 5) -> blur 64 -> sRGB",
                   "l2": "every pass streams 256 MiB layers (> 126 MB L2)", "sharding": "one layer per GPU, no collective"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_box_blur_v2 + k_box_blur_h2 (20 of the chain's passes)",
                     "achieved": blur_bytes / (ms_blur * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": blur_bytes / (ms_blur * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes": blur_bytes, "kernel_ms": ms_blur, "model": "8 B/px per box pass, 20 passes"},
        "e2e": {"value": shard.aggregate_throughput(mpx, world, e2e_s), "unit": "Mpx/s", "h2d_bytes_per_step": W * H * 4,
                "d2h_bytes_per_step": W * H * 4, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        from tests import oracle_ffi as O  # the CPU checker: cpu_baseline only
        n = 1024
        img = np.ascontiguousarray(host_src[:n, :n])
        t0 = time.perf_counter()
        x = O.into_linear_rgb(img)
        x = O.box_blur(8.0, 8.0, x)
        x = O.morphology("dilate", 3.0, 3.0, x)
        x = O.convolve_matrix(sharpen, 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, x)
        tb = O.multiply_alpha(O.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, 3, 7, False, False, n, n))
        x = O.arithmetic(0.5, 0.5, 0.5, 0.0, x, tb)
        x = O.diffuse_lighting(5.0, 1.0, (255, 255, 255), O.make_light("distant", azimuth=45.0, elevation=60.0), x)
        x = O.into_srgb(O.box_blur(64.0, 64.0, x))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n * n / 1e6 / dt, "unit": "Mpx/s", "cores": 1, "kind": "port",
                               "sample": f"the same chain on the top-left {n}x{n} px of the source layer, {dt:.2f} s; oracle restatement"}
        # parity: the GPU chain over the same crop (a layer of its own) against the checker's result
        a, t, c = ctx.layer_from(img), ctx.layer(n, n), ctx.layer(n, n)
        chain()
        out["parity"] = parity_record(a.download(), x, f"the chain over the top-left {n}x{n} crop of the source layer, GPU vs CPU checker")
    return out


def stack_tree_stream(name):
    """The usvg tree of a stack document as the RBT1 stream the Rust shim would write (tools/make_stack_trees.py)."""
    with open(os.path.join(ROOT, "resvg_b200", "data", name), "rb") as f:
        return f.read()


def run_stack4k(args, rank, local_rank, world, torch, dist):
    """Workload `stack4k` (BASELINE configs[3]): 64 nested groups with opacity, luminance masks, clip-paths
    (every 8th nested) and pattern fills on a 4096x4096 canvas.  The document reaches the library as a usvg tree stream
    (parsing SVG is usvg's job and stays on the host; the streams are committed, see stack_tree_stream) and ONE rb_render
    call draws it: traversal (render.rs / clip.rs / mask.rs / path.rs), layers and every pixel inside libresvg_b200.so.
    value: CUDA-event time of rb_render, tree resident; e2e: rb_submit from the host stream (parse + render) + download."""
    import resvg_b200 as rb
    from resvg_b200 import shard
    W, H, levels, seed = WORKLOADS["stack4k"]
    ctx = rb.Context(local_rank)
    # --shard strips: ONE document, rank r renders rows strip_for_rank(H, r, world) of it (rb_render_strip:
    # bit-identical to the whole render, tests/test_stack.py); otherwise one document per GPU
    strips = args.shard == "strips"
    y0, rows = shard.strip_for_rank(H, rank, world) if strips else (0, H)
    blob = stack_tree_stream(f"stack4k_r{0 if strips else rank % 8}.rbt")
    tree = rb.tree.Tree(blob)
    ident = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    target = ctx.layer(W, rows)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def step():
        target.fill(0, 0, 0, 0)
        if strips:
            rb.tree.render_strip(tree, ident, W, H, y0, target)
        else:
            rb.tree.render(tree, ident, target)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms_step = ctx.timer_end() / args.steps
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0

    # the traversal's most frequent full-layer kernel alone: the layer composite (draw_pixmap), 12 B/px
    a, b2 = ctx.layer(W, H), ctx.layer(W, H)
    rb.draw_layer(a, b2, 0, 0, 0.9, "source_over")
    ctx.timer_begin()
    for _ in range(10):
        rb.draw_layer(a, b2, 0, 0, 0.9, "source_over")
    ms_comp = ctx.timer_end() / 10
    a.close(); b2.close()

    pinned = rb.PinnedBuffer(W * rows * 4)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))

    def e2e_step():
        target.fill(0, 0, 0, 0)
        if strips:
            t = rb.tree.Tree(blob)                   # host stream -> parse
            rb.tree.render_strip(t, ident, W, H, y0, target)
            t.close()
        else:
            rb.tree.submit(blob, ident, target)      # host stream -> parse -> traversal -> kernels
        target.download_ptr(pinned.array.ctypes.data)

    e2e_step()
    barrier()
    h2d0 = ctx.h2d_bytes
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = (ctx.h2d_bytes - h2d0) // e2e_steps
    ms_step, ms_comp, e2e_s = shard.max_over_ranks([ms_step, ms_comp, e2e_s], world, f"cuda:{local_rank}")
    if rank != 0:
        return None
    mpx = W * H / 1e6
    peak, peak_src = measured_peaks()
    comp_bytes = 12 * W * H
    out = {
        "metric": "Mpixels/s rendered", "value": shard.aggregate_throughput(mpx, 1 if strips else world, ms_step * 1e-3), "unit": "Mpx/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strips else "weak", "vs_baseline": None, "dtype": "u8/u16 fixed point + f32", "data": "synthetic",
        "config": {"workload": "stack4k", "canvas": [W, H], "levels": levels,
                   "recipe": "level k: opacity U[0.85,0.99]; k%4=0 luminance mask (gradient rect), 1 clip-path (circle, every 8th nested), "
                             "2 pattern-filled rect (tile 32-128 px), 3 plain; 10 C2-style shapes per level in a box inset 16 px per level",
                   "host": "the usvg tree arrives as an RBT1 stream (%d bytes); ONE rb_render call per document, traversal in C++ inside the library" % len(blob),
                   "l2": "layers of up to 64 MiB; every level allocates, composites and masks its layer (> 126 MB L2 across a level)",
                   "sharding": ("ONE document cut into %d canvas strips (rb_render_strip): draws that reach the canvas directly are built against "
                                "the whole canvas, isolated groups are rendered as in the whole render and composited shifted, groups that miss "
                                "the strip are skipped; no halo, no collective, every rank downloads its strip" % world) if strips
                               else "one document per GPU, no collective"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_draw_layer (layer composite, the traversal's most frequent full-layer pass)",
                     "achieved": comp_bytes / (ms_comp * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": comp_bytes / (ms_comp * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes": comp_bytes, "kernel_ms": ms_comp, "model": "12 B/px: source read, destination read + write"},
        "e2e": {"value": shard.aggregate_throughput(mpx, 1 if strips else world, e2e_s), "unit": "Mpx/s", "h2d_bytes_per_step": int(h2d) + len(blob),
                "d2h_bytes_per_step": W * rows * 4, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        from tests import svgfront as F               # the CPU checker's traversal: cpu_baseline / parity only
        from tests.backends import OracleBackend
        from tests import scenes_loader
        n = 1024
        small = F.parse(scenes_loader.load().stack_svg(n, levels, seed))
        t0 = time.perf_counter()
        ref = F.render_scene(small, OracleBackend(), n)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n * n / 1e6 / dt, "unit": "Mpx/s", "cores": 1, "kind": "port",
                               "sample": f"the same recipe at {n}x{n} (1/16 of the pixels, same 64 levels) through the checker's traversal, {dt:.2f} s; "
                                         "oracle restatement of the resvg/tiny-skia CPU path"}
        small_l = ctx.layer(n, n)
        rb.tree.submit(stack_tree_stream("stack1k.rbt"), ident, small_l)
        out["parity"] = parity_record(small_l.download(), ref, f"the same {n}x{n} document: rb_submit vs the CPU checker (f32 composites: <= 1)")
    return out


def run_reference_icons(args, cores):
    """--impl reference --workload icons: the CPU checker renders documents one at a time on every host core (one document
    stream per thread, as `resvg` would be run per file)."""
    from tests import icons_ref
    scenes = icons_ref.scenes
    n_docs = int(args.docs) if args.docs else WORKLOADS["icons"][2]
    per = max(8, min(256, int(args.cpu_sample) // (8 * cores)))
    sc = scenes.icons_docs(0, per * cores)
    paints = icons_ref.prepare(sc)

    def step():
        ths = [threading.Thread(target=icons_ref.render_docs, args=(sc, paints, range(t * per, (t + 1) * per))) for t in range(cores)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    dt = statistics.mean(times)
    v = per * cores * 0.065536 / dt
    sample = f"{cores} threads x {per} documents (documents 0..{per * cores - 1} of the batch) per step, one document at a time"
    return {
        "impl": "reference", "metric": "Mpixels/s rendered", "value": v, "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16 fixed point + f32", "data": "synthetic",
        "config": {"workload": "icons", "documents": n_docs, "document_px": [256, 256]},
        "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on all host cores.  The Rust reference cannot be built here, so this is the
    oracle restatement ("port") — pixels by oracle/raster.c + filters.c, stroking / dashing / hairlines by oracle/stroke.c,
    dash.c, hairline.c.  Nothing of the product is imported: the only native library this arm loads is oracle/liboracle.so."""
    if rank != 0:
        return None
    from tests import scenes_loader
    scenes = scenes_loader.load()
    cores = min(os.cpu_count() or 1, 32)
    if args.workload == "icons":
        return run_reference_icons(args, cores)
    if args.workload in ("filters8k", "stack4k"):
        raise SystemExit("--impl reference: workloads filters8k / stack4k have a cpu_baseline in the ours arm only")
    R = oracle_lib()
    W, H, n_paths, seed = WORKLOADS[args.workload]
    scs = [scenes.paths_scene(W, H, n_paths, seed + t) for t in range(min(cores, 4))]
    n_draws = scs[0]["n_paths"]
    # sized so that warmup + steps fit a few minutes: every thread renders the first n_sample draws on a full canvas
    budget = max(1, args.steps + args.warmup)
    n_sample = max(200, min(n_draws, int(args.cpu_sample), int(args.cpu_sample) * 13 // budget))
    canvases = [np.zeros((H, W, 4), np.uint8) for _ in range(cores)]
    # Host geometry (dash / stroke / hairline walk) is prepared once, outside the timed region, because the Python loop
    # around those C calls would serialise the threads on the GIL; the seconds spent INSIDE the C calls are added to
    # every step (each thread would spend them in parallel), so the reference is not charged for Python overhead.
    preps = [prepare_sample(R, sc, n_sample) for sc in scs]
    t_geom = statistics.mean(p["t_geom"] for p in preps)

    def step():
        for c in canvases:
            c[...] = 0
        ths = [threading.Thread(target=render_prepared, args=(R, preps[t % len(preps)], canvases[t])) for t in range(cores)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0 + t_geom

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    dt = statistics.mean(times)
    mpx = cores * (W * H / 1e6) * (n_sample / n_draws)
    v = mpx / dt
    sample = (f"{cores} threads x first {n_sample} of {n_draws} draws of the {W}x{H} scene per step (full canvas); value scaled by "
              f"{n_sample}/{n_draws}; step = threaded oracle pixel work (wall) + {t_geom:.2f} s of oracle stroking/dashing/hairline walking per thread")
    return {
        "impl": "reference", "metric": "Mpixels/s rendered", "value": v, "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16 fixed point + f32", "data": "synthetic",
        "config": {"workload": args.workload, "canvas": [W, H], "paths": n_paths, "draw_calls": int(n_draws)},
        "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="paths8k", choices=list(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=12000, help="paths in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true", help="skip the per-kernel roofline table of the filter / compositing kernels")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--shard", default="documents", choices=["documents", "strips"],
                    help="paths8k at N > 1: one scene per GPU (weak scaling, default) or ONE scene cut into N canvas strips (strong)")
    ap.add_argument("--docs", type=int, default=0, help="icons: number of documents (default: the 100 000 of the recipe)")
    ap.add_argument("--no-configs", action="store_true",
                    help="paths8k: skip the sub-records of the other BASELINE configurations (icons, filters8k, stack4k)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: resvg_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload in ("icons", "filters8k", "stack4k"):
        out = {"icons": run_icons, "filters8k": run_filters8k, "stack4k": run_stack4k}[args.workload](args, rank, local_rank, world, torch, dist)
        if out is not None:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    import resvg_b200 as rb
    from resvg_b200 import _ffi, scenes, shard

    W, H, n_paths, seed = WORKLOADS[args.workload]
    canvas_mpx = W * H / 1e6
    ctx = rb.Context(local_rank)
    strips = args.shard == "strips"
    y0, strip_rows = shard.strip_for_rank(H, rank, world) if strips else (0, H)
    draw_vp = shard.strip_viewport(W, H, y0) if strips else None
    # every rank renders its own document — or, with --shard strips, rows y0 .. y0 + strip_rows of the same document
    scene = scenes.paths_scene(W, H, n_paths, seed if strips else shard.scene_seed(seed, rank))
    n_threads = 0 if world == 1 else shard.host_threads(world)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    n_draws_in = scene["n_paths"]

    layer = ctx.layer(W, strip_rows)
    batch = rb.Batch(layer)
    if draw_vp:
        batch.set_viewport(*draw_vp)
    batch.fill_paths(scene)
    batch.prepare(n_threads)  # host edge build + binning + H2D: inputs are resident in HBM before the timed region
    st = batch.stats()
    ctx.synchronize()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def step():
        layer.fill(0, 0, 0, 0)
        batch.run()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms_total = ctx.timer_end()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0
    ms_step = ms_total / args.steps

    # dominant kernel alone (k_raster_warp), for the roofline: CUDA events around that launch inside the library
    ms_kernel, ms_prepass = 0.0, 0.0
    for _ in range(args.steps):
        layer.fill(0, 0, 0, 0)
        batch.run()
        pre, ras = ctx.last_run_ms()
        ms_prepass += pre / args.steps
        ms_kernel += ras / args.steps
    layer.fill(0, 0, 0, 0)
    n_rmw, n_store = batch.run_counting()
    alg_bytes = 8 * n_rmw + 4 * n_store + 16 * st["edges"]

    # end to end through the C ABI with host buffers: record + edge build + H2D + kernel + D2H, every step
    pinned = rb.PinnedBuffer(W * strip_rows * 4)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_h2d = 0

    banded_download = os.environ.get("RB_BENCH_PLAIN_DOWNLOAD") != "1"
    e2e_phase = [0.0, 0.0, 0.0]  # seconds: record, host build + enqueue, wait for the GPU + D2H
    e2e_geo = [0] * 6  # rb_debug_geo_counts after the last step

    def e2e_step():
        nonlocal e2e_h2d
        ta = time.perf_counter()
        layer.fill(0, 0, 0, 0)
        b = rb.Batch(layer)
        if draw_vp:
            b.set_viewport(*draw_vp)
        b.fill_paths(scene)
        tb = time.perf_counter()
        if banded_download:
            b.submit_download(pinned.array.ctypes.data, n_threads)  # the download overlaps the last raster launch, band by band
        else:
            b.submit(n_threads)
        tc = time.perf_counter()
        e2e_h2d = b.stats()["upload_bytes"]
        gc = (C.c_uint64 * 6)()
        _ffi.lib.rb_debug_geo_counts(gc)
        e2e_geo[:] = [int(v) for v in gc]
        b.close()
        if banded_download:
            layer.download_end()  # synchronises
        else:
            layer.download_ptr(pinned.array.ctypes.data)  # synchronises
        td = time.perf_counter()
        e2e_phase[0] += tb - ta; e2e_phase[1] += tc - tb; e2e_phase[2] += td - tc

    e2e_step()
    barrier()
    e2e_phase[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    per_rank = shard.gather_floats([e2e_s * 1e3, e2e_phase[1] / e2e_steps * 1e3, e2e_phase[2] / e2e_steps * 1e3, e2e_geo[3] / 1e3], world, f"cuda:{local_rank}")
    ms_step, ms_kernel, e2e_s = shard.max_over_ranks([ms_step, ms_kernel, e2e_s], world, f"cuda:{local_rank}")

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get("k_raster_warp_dram_bytes")
        except Exception:
            pass
        out = {
            "metric": "Mpixels/s rendered", "value": shard.aggregate_throughput(canvas_mpx, 1 if strips else world, ms_step * 1e-3), "unit": "Mpx/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if strips else "weak", "vs_baseline": None,
            "dtype": "u8/u16 fixed point + f32", "data": "synthetic",
            "config": {"workload": args.workload, "canvas": [W, H], "paths": n_paths, "draw_calls": int(n_draws_in),
                       "mix": "70% fill / 20% fill+stroke / 10% stroke; nonzero+evenodd; 50% solid / 30% linear / 20% radial; 95% AA",
                       "stroke": "width log-U[0.5,16] px (<= 1 px: anti-aliased hairlines), miter/round/bevel joins, butt/round/square caps, 10% dashed (2-4 intervals U[2,32])",
                       "sharding": ("ONE scene cut into %d canvas strips of %d rows: every rank records the whole scene with the document's "
                                    "pixmap placed above its strip layer (rb_batch_set_viewport with a negative origin), so the strips hold "
                                    "exactly the pixels of the whole-canvas render; no collective, every rank downloads its strip"
                                    % (world, strip_rows)) if strips
                                   else "one scene (document) per GPU, no collective",
                       "l2": "inputs (256 MiB canvas + %.0f MiB edges/bins) exceed the 126 MB L2" % (st["upload_bytes"] / 2**20),
                       "draws": st["draws"], "line_edges": st["edges"], "draw_tile_pairs": st["pairs"], "tiles": st["tiles"]},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_raster_warp", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes, "kernel_ms": ms_kernel, "prepass_ms": ms_prepass,
                         "model": "8 B per blended px + 4 B per opaque-stored px + 16 B per line edge"},
            "e2e": {"value": shard.aggregate_throughput(canvas_mpx, 1 if strips else world, e2e_s), "unit": "Mpx/s", "h2d_bytes_per_step": int(e2e_h2d),
                    "d2h_bytes_per_step": W * strip_rows * 4, "ms_per_step": e2e_s * 1e3, "record_ms": e2e_phase[0] / e2e_steps * 1e3,
                    "host_build_and_enqueue_ms": e2e_phase[1] / e2e_steps * 1e3, "gpu_wait_and_d2h_ms": e2e_phase[2] / e2e_steps * 1e3,
                    "steps": e2e_steps,
                    "download": ("rb_batch_submit_download: the last raster launch runs in bands of tile rows, each copied out while the next "
                                 "is rendered; rb_layer_download_end waits") if banded_download
                                else "rb_batch_submit, then rb_layer_download",
                    "per_rank_ms": {"e2e": [round(r[0], 2) for r in per_rank], "host_build_and_enqueue": [round(r[1], 2) for r in per_rank],
                                    "gpu_wait_and_d2h": [round(r[2], 2) for r in per_rank], "device_geometry": [round(r[3], 2) for r in per_rank]},
                    "geometry": {"where": "device" if e2e_geo[0] > 0 and e2e_geo[1] == 0 else "host",
                                 "ranges_built_on_device": e2e_geo[0], "ranges_handed_to_host_builder": e2e_geo[1], "heap_retries": e2e_geo[2],
                                 "device_geometry_ms": e2e_geo[3] / 1e3, "host_task_build_ms": e2e_geo[4] / 1e3, "tasks": e2e_geo[5],
                                 "note": "dash / stroke / hairline walk / chop / clip / edge set-up run as CUDA kernels (geo.cu) from the raw "
                                         "paths; host_build_and_enqueue_ms includes waiting for them (their totals size the raster scratch)"}},
        }
        if world == 1 and not args.no_kernel_table:
            out["kernels"] = kernel_table(rb, ctx, layer, W, H, peak)
        if world == 1 and not args.no_cpu_baseline:
            R = oracle_lib()
            n_sample = max(200, min(n_draws_in, int(args.cpu_sample)))
            ref_canvas = np.zeros((H, W, 4), np.uint8)
            dt = cpu_render_sample(R, scene, n_sample, ref_canvas)
            out["cpu_baseline"] = {
                "value": canvas_mpx / (dt * n_draws_in / n_sample), "unit": "Mpx/s", "cores": 1, "kind": "port",
                "sample": f"first {n_sample} of {n_draws_in} draws of the same scene on the full canvas, {dt:.2f} s; "
                          f"value = canvas Mpx / (t * {n_draws_in}/{n_sample}); oracle restatement of the resvg/tiny-skia CPU path "
                          "(pixels, stroker, dasher, hairline walker all oracle/)"}
            # parity at BASELINE size: the same first n_sample draws rendered by the product through the C ABI, against the
            # canvas the checker has just produced (each arm with its OWN stroker / dasher / hairline walker)
            sub = scenes.subset(scene, n_sample)
            sub["paints"] = scenes.to_paint_array(sub, _ffi.Paint)
            sub["strokes"] = scenes.to_stroke_array(sub, _ffi.Stroke)
            layer.fill(0, 0, 0, 0)
            pb = rb.Batch(layer)
            pb.fill_paths(sub)
            pb.submit(n_threads)
            pb.close()
            out["parity"] = parity_record(layer.download(), ref_canvas,
                                          f"first {n_sample} draws on the full {W}x{H} canvas: GPU (C ABI) vs CPU checker")
            del ref_canvas
        if world == 1 and not args.no_configs and args.workload == "paths8k":
            # the other BASELINE configurations, each with its own value / e2e / roofline / cpu_baseline / parity
            # (full recipes except icons, which renders 8 atlas chunks = 8192 documents here; `--workload icons` runs all 100 000)
            batch.close()
            layer.close()
            sub_args = argparse.Namespace(**vars(args))
            sub_args.steps, sub_args.warmup, sub_args.e2e_steps = max(3, min(args.steps, 5)), 3, 2
            sub_args.cpu_sample = min(int(args.cpu_sample), 2048)
            configs = {}
            sub_args.docs = 8192
            configs["icons"] = run_icons(sub_args, rank, local_rank, world, torch, dist)
            sub_args.docs = 0
            configs["filters8k"] = run_filters8k(sub_args, rank, local_rank, world, torch, dist)
            configs["stack4k"] = run_stack4k(sub_args, rank, local_rank, world, torch, dist)
            out["configs"] = configs
            out["gpu_launches_note"] = "gpu_launches counts the timed paths8k steps only; every sub-record carries its own"
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
