#!/bin/bash
# compute-sanitizer passes over the GPU tests (run through gpurun on a B200).
set -u
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_raster_gpu.py tests/test_edge_cases_gpu.py -q -x -k "not full_size" || exit 1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_filters_gpu.py -q -x -k "box or iir or helpers or morph or convolve" || exit 1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_raster_gpu.py -q -x \
    -k "structured or hairline or batch_painters or coverage_random_paths or viewports or masks" || exit 1
# the atlas path (region composites, cell-aware blur, flood) and the remaining filter kernels
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_icons.py tests/test_stack.py -q -x -m gpu -k "not chunk_or_cell" || exit 1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_filters_gpu.py -q -x -k "not (box or iir or helpers or morph or convolve)" || exit 1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_filters_gpu.py tests/test_icons.py -q -x -m gpu -k "box or morph or convolve or cells" || exit 1
# the geometry kernels (geo.cu): every case but the 8192 x 8192 ones, and the raster tests with every batch on the device
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_geo_gpu.py -q -x -k "not full_size and not draw_tiler" || exit 1
RB_GEO_MODE=1 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_raster_gpu.py -q -x -k "hairline or dashed or viewports or bench_scene_bulk" || exit 1
# the dealt-out gradient blend (shared-memory staging in k_raster_warp) and the banded download
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_raster_gpu.py -q -x -k "test_gradients or radial_and_focal or bench_scene_bulk" || exit 1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_geo_gpu.py -q -x -k "submit_download and 1500" || exit 1
