"""Document-parallel sharding of the pixel path across GPUs.

resvg renders one tree into one pixmap (crates/resvg/src/lib.rs:34); independent documents share nothing, so the
multi-GPU form of the path is one process per GPU, each rendering its own documents, with NO collective on the data
path.  torch.distributed is used for the timing protocol only (barrier, max-over-ranks), which is what this module
holds so that it can be exercised on CPU with the gloo backend (tests/test_shard_gloo.py) and on GPUs with nccl
(bench.py).
"""
from __future__ import annotations

import os


def env_rank():
    """(rank, local_rank, world) as torchrun exports them; (0, 0, 1) for a plain launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def documents_for_rank(n_documents: int, rank: int, world: int):
    """Round-robin assignment of document indices: rank r renders documents r, r + world, ..."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_documents, world))


def scene_seed(base_seed: int, rank: int, index: int = 0, world: int = 1) -> int:
    """Seed of the synthetic document `index` of `rank` (distinct per rank so ranks do not render one scene)."""
    return int(base_seed) + rank + index * world


def host_threads(world: int) -> int:
    """Host edge-build threads per rank: the ranks of one node share its cores."""
    return max(1, (os.cpu_count() or 1) // max(1, world))


def max_over_ranks(values, world: int, device=None):
    """Element-wise maximum over ranks of a small list of floats (device timings); identity for world == 1."""
    if world == 1:
        return [float(v) for v in values]
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_ints(values, world: int, device=None):
    """All ranks' integer tuples (checksums, counters), as a list indexed by rank."""
    if world == 1:
        return [[int(v) for v in values]]
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[int(v) for v in o.tolist()] for o in out]


def gather_floats(values, world: int, device=None):
    """All ranks' float tuples (per-rank timings), as a list indexed by rank."""
    if world == 1:
        return [[float(v) for v in values]]
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[float(v) for v in o.tolist()] for o in out]


def aggregate_throughput(units_per_rank: float, world: int, seconds: float) -> float:
    """Whole-job throughput under weak scaling: every rank processed `units_per_rank` in `seconds` (max over ranks)."""
    return world * units_per_rank / seconds


def strip_for_rank(height: int, rank: int, world: int, align: int = 8):
    """Canvas-strip sharding of ONE large canvas: rank r renders rows [y0, y0 + rows) of the
    canvas into a layer of its own; strips are `align`-row aligned (the raster kernel's tile height), cover [0, height)
    exactly once and differ in size by at most one aligned block.  Path rendering needs no halo and no exchange: every
    rank records the whole scene with `strip_viewport`, i.e. the document's pixmap placed y0 rows above its strip layer,
    so every path is clipped and flattened against the WHOLE canvas and the strip receives exactly the pixels of the
    whole-canvas render."""
    if world <= 0 or not (0 <= rank < world) or height <= 0:
        raise ValueError("bad rank/world/height")
    blocks = (height + align - 1) // align
    lo = (blocks * rank) // world
    hi = (blocks * (rank + 1)) // world
    y0, y1 = lo * align, min(hi * align, height)
    return y0, max(0, y1 - y0)


def strip_viewport(width: int, height: int, y0: int):
    """rb_batch_set_viewport arguments that place the width x height document so that its row y0 is row 0 of the strip layer."""
    return (0, -int(y0), int(width), int(height))


def strip_transform(y0: int):
    """The DrawTiler-style alternative: translate the draws by -y0 and let the strip layer be the pixmap.  Curves crossing
    a strip boundary are then clipped to the strip before flattening and differ slightly from the whole-canvas render;
    kept for the test that documents the difference."""
    return (1.0, 0.0, 0.0, 1.0, 0.0, -float(y0))


def strip_with_halo(height: int, rank: int, world: int, halo: int, align: int = 8):
    """Rows a rank must hold to FILTER its strip of a large layer: its own rows plus `halo` rows of
    context on either side, clipped to the image — the sum of the vertical reach of the chain's primitives (box blur:
    filters.box_blur_reach(sigma); morphology: ceil(ry); convolve: rows - 1; lighting: 1; pointwise: 0).  Rendering
    the halo redundantly replaces any exchange between the GPUs.  Returns (first_row, n_rows, offset of the strip's own first
    row inside that block, own rows)."""
    y0, rows = strip_for_rank(height, rank, world, align)
    lo = max(0, y0 - halo)
    hi = min(height, y0 + rows + halo)
    return lo, hi - lo, y0 - lo, rows
