"""Worker of tests/test_shard_gloo.py: one rank of the document-parallel path on CPU (gloo).

Each rank builds ITS document (resvg_b200.shard.scene_seed), renders it with the oracle (there is no GPU here; on the
GPU box bench.py runs the same protocol with the CUDA path and nccl), then the ranks exchange checksums and reduce
their timings exactly as bench.py does.  Rank 0 writes a JSON report.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, W, H, n_paths, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    import torch.distributed as dist
    import bench
    from resvg_b200 import scenes, shard

    rank, _, world = shard.env_rank()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    try:
        scene = scenes.paths_scene(W, H, n_paths, shard.scene_seed(seed, rank))
        px = np.zeros((H, W, 4), np.uint8)
        dt = bench.cpu_render_sample(bench.oracle_lib(), scene, scene["n_paths"], px)
        crc = zlib.crc32(px.tobytes())
        dist.barrier()
        per_rank = shard.gather_ints([crc, scene["n_paths"], int(px[..., 3].astype(np.int64).sum())], world)
        fake_ms = 10.0 + 5.0 * rank  # deterministic "device times": the reduction must return the slowest rank's
        (ms,) = shard.max_over_ranks([fake_ms], world)
        (real,) = shard.max_over_ranks([dt], world)
        docs = shard.documents_for_rank(7, rank, world)
        all_docs = shard.gather_ints(docs + [-1] * (4 - len(docs)), world)
        # the regression corpus sharded by file: fixture i -> rank i mod world, every rank diffs its own
        # files against the reference goldens; only counts and a checksum cross ranks
        import glob
        from PIL import Image
        from tests import svgfront as F
        from tests.backends import OracleBackend
        files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scenes", "*.json")))[:16]
        passed, ccrc = 0, 0
        for i in shard.documents_for_rank(len(files), rank, world):
            with open(files[i]) as f:
                sc = json.load(f)
            img = F.render_scene(sc, OracleBackend(), 300)
            ccrc = zlib.crc32(img.tobytes(), ccrc)
            passed += F.diff_pixels(img, np.array(Image.open(files[i][:-5] + ".png").convert("RGBA"))) == 0
        corpus = shard.gather_ints([passed, ccrc], world)
        # one document cut into canvas strips: rank r owns rows strip_for_rank(height, r, world) of the render
        # (on a GPU: rb_render_strip; here the checker renders the document and keeps the rank's rows); the strips tile the
        # canvas and only their checksums cross ranks
        from resvg_b200 import scenes as _scenes
        from tests import svgfront as _F
        doc = _F.parse(_scenes.stack_svg(96, 6, inset=3.0, shapes=3))
        whole = _F.render_scene(doc, OracleBackend(), 96)
        y0, rows = shard.strip_for_rank(whole.shape[0], rank, world)
        strips = shard.gather_ints([y0, rows, zlib.crc32(np.ascontiguousarray(whole[y0:y0 + rows]).tobytes())], world)
        if rank == 0:
            with open(out_path, "w") as f:
                json.dump({"world": world, "per_rank": per_rank, "ms": ms, "render_s": real, "docs": all_docs, "corpus": corpus, "strips": strips,
                           "value": shard.aggregate_throughput(W * H / 1e6, world, ms * 1e-3)}, f)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
