#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: raster.cu compiled with extra defines, linked into build_variants/<name>.so
name=$1; shift
mkdir -p build_variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v --expt-relaxed-constexpr -Wno-deprecated-gpu-targets"
$NV "$@" -c resvg_b200/csrc/raster.cu -o build_variants/raster_$name.o 2> build_variants/raster_$name.ptxas.log || { cat build_variants/raster_$name.ptxas.log; exit 1; }
objs=$(ls resvg_b200/csrc/*.o | grep -v "/raster.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/$name.so $objs build_variants/raster_$name.o -lpthread
grep -A2 "k_raster_warpILb0ELb1ELb0" build_variants/raster_$name.ptxas.log | grep -E "stack|registers"
