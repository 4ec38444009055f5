"""World-size-2 CPU (gloo) run of the multi-GPU protocol: document-parallel sharding, no data-path collective
(SURVEY.md §8(e)); only checksums and timings cross ranks."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from resvg_b200 import scenes, shard  # noqa: E402


def test_document_assignment_is_a_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(d for r in range(world) for d in shard.documents_for_rank(11, r, world))
        assert seen == list(range(11))
    with pytest.raises(ValueError):
        shard.documents_for_rank(4, 2, 2)
    assert shard.scene_seed(7, 0) != shard.scene_seed(7, 1)
    assert len({shard.scene_seed(7, r, i, 4) for r in range(4) for i in range(5)}) == 20
    assert shard.aggregate_throughput(67.108864, 8, 0.05) == pytest.approx(8 * 67.108864 / 0.05)
    assert shard.max_over_ranks([1.5, 2.5], 1) == [1.5, 2.5]
    assert shard.host_threads(1) >= shard.host_threads(2) >= 1


def test_corpus_files_are_partitioned_by_file():
    """C1 (the regression corpus) shards by file: file i -> rank i mod N; every fixture is rendered exactly once."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scenes", "*.json")))
    assert len(files) >= 850
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in shard.documents_for_rank(len(files), r, world))
        assert got == list(range(len(files)))
        sizes = [len(shard.documents_for_rank(len(files), r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_strips_partition_the_canvas():
    for height in (8192, 4096, 300, 17, 8):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                y0, n = shard.strip_for_rank(height, r, world)
                assert y0 % 8 == 0 and n >= 0
                rows += list(range(y0, y0 + n))
            assert rows == list(range(height)), (height, world)
    y0, n = shard.strip_for_rank(8192, 3, 8)
    assert (y0, n) == (3072, 1024) and shard.strip_viewport(8192, 8192, y0) == (0, -3072, 8192, 8192)
    with pytest.raises(ValueError):
        shard.strip_for_rank(100, 2, 2)
    # strips with a filter halo: own rows unchanged, context clipped to the image
    assert shard.strip_with_halo(1000, 0, 2, 40) == (0, 536, 0, 496)
    assert shard.strip_with_halo(1000, 1, 2, 40) == (456, 544, 40, 504)


@pytest.mark.timeout(300)
def test_two_ranks_gloo(tmp_path):
    import bench
    W, H, n_paths, seed = 192, 160, 60, 4242
    out = tmp_path / "report.json"
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(out), str(W), str(H), str(n_paths), str(seed)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(out.read_text())
    assert rep["world"] == 2
    # every rank rendered its own document: same result as rendering the two documents serially here
    R = bench.oracle_lib()
    want = []
    for rank in range(2):
        scene = scenes.paths_scene(W, H, n_paths, shard.scene_seed(seed, rank))
        px = np.zeros((H, W, 4), np.uint8)
        bench.cpu_render_sample(R, scene, scene["n_paths"], px)
        want.append([zlib.crc32(px.tobytes()), scene["n_paths"], int(px[..., 3].astype(np.int64).sum())])
    assert rep["per_rank"] == want
    assert want[0][0] != want[1][0]
    assert rep["ms"] == 15.0  # max over ranks, not rank 0's 10 ms
    assert rep["value"] == pytest.approx(2 * (W * H / 1e6) / 15e-3)
    assert rep["docs"] == [[0, 2, 4, 6], [1, 3, 5, -1]]
    # corpus by file: 16 fixtures over 2 ranks, all reproduce their golden, and the per-rank checksums equal a serial run's
    import glob
    import json as _json
    from tests import svgfront as F
    from tests.backends import OracleBackend
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scenes", "*.json")))[:16]
    serial = []
    for r in range(2):
        crc = 0
        for i in range(r, 16, 2):
            with open(files[i]) as f:
                crc = zlib.crc32(F.render_scene(_json.load(f), OracleBackend(), 300).tobytes(), crc)
        serial.append([8, crc])
    assert rep["corpus"] == serial
    # one document in canvas strips: the ranks' rows tile the canvas, their checksums are those of the whole render's rows
    doc = F.parse(scenes.stack_svg(96, 6, inset=3.0, shapes=3))
    whole = F.render_scene(doc, OracleBackend(), 96)
    want_strips = []
    for r in range(2):
        y0, rows = shard.strip_for_rank(whole.shape[0], r, 2)
        want_strips.append([y0, rows, zlib.crc32(np.ascontiguousarray(whole[y0:y0 + rows]).tobytes())])
    assert rep["strips"] == want_strips
    assert want_strips[0][0] == 0 and want_strips[0][1] == want_strips[1][0] and want_strips[1][0] + want_strips[1][1] == whole.shape[0]
