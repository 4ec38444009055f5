#!/usr/bin/env python
"""BASELINE configs[0] (the resvg regression corpus) sharded BY FILE across GPUs:

    python tools/corpus_gpu.py                       # one GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/corpus_gpu.py

File i of the sorted fixture list goes to rank i mod N (shard.documents_for_rank); every rank renders its files with ONE
rb_render call each and diffs them against the reference's golden PNG at the reference's own criterion
(tests/integration/main.rs:151-226).  No data-path collective: only the pass / fail counts and a checksum of the rendered
bytes are gathered.  Prints one JSON line on rank 0."""
import glob
import json
import os
import sys
import time
import zlib

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resvg_b200 as rb  # noqa: E402
from resvg_b200 import shard  # noqa: E402
from tests import svgfront as F  # noqa: E402  (diff criterion + target size only; no rendering)


def main():
    rank, local_rank, world = shard.env_rank()
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scenes", "*.json")))
    mine = shard.documents_for_rank(len(files), rank, world)
    ctx = rb.Context(local_rank)
    scenes = []
    for i in mine:
        with open(files[i]) as f:
            scenes.append((files[i], json.load(f), np.array(Image.open(files[i][:-5] + ".png").convert("RGBA"))))
    trees = [rb.tree.Tree(sc) for _, sc, _ in scenes]  # uploaded once, like resvg_parse_tree
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    passed, failed, crc, px = 0, [], 0, 0
    for (path, sc, gold), tree in zip(scenes, trees):
        pw, ph, ts = F.target_for(sc, 300)
        layer = ctx.layer(pw, ph)
        rb.tree.render(tree, ts, layer)
        out = layer.download()
        layer.close()
        px += pw * ph
        crc = zlib.crc32(out.tobytes(), crc)
        if F.diff_pixels(out, gold) == 0:
            passed += 1
        else:
            failed.append(os.path.basename(path)[:-5])
    ctx.synchronize()
    dt = time.perf_counter() - t0
    rows = shard.gather_ints([passed, len(failed), crc, px, int(dt * 1e6)], world, f"cuda:{local_rank}")
    if rank == 0:
        tot_pass, tot_fail = sum(r[0] for r in rows), sum(r[1] for r in rows)
        print(json.dumps({"workload": "resvg regression corpus, sharded by file", "n_gpus": world, "files": len(files), "pass": tot_pass,
                          "fail": tot_fail, "failed_on_rank0": failed, "per_rank": [{"files": r[0] + r[1], "crc32": r[2], "px": r[3], "s": r[4] / 1e6} for r in rows],
                          "Mpx_per_s_wall": sum(r[3] for r in rows) / max(r[4] for r in rows),
                          "note": "wall time per rank includes the PNG-free diff on the host; one rb_render call per file"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
