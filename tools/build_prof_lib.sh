#!/bin/sh
# Measurement build of the library: the warp-per-chain TVP kernel with per-phase cycle counters (BVG_TVP_PROF), linked with
# the product objects into build_alt/libbvargpu_prof.so (git-ignored).  Use: BVG_LIB=build_alt/libbvargpu_prof.so python tools/micro/tvp_prof.py
set -e
cd "$(dirname "$0")/.."
make -s -j8 -C bvartools_b200/csrc
mkdir -p build_alt
cd bvartools_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DBVG_G=32 -DBVG_TVP_PROF \
     -c bvg_kern_tvp.cu -o ../../build_alt/bvg_kern_tvp_g32_prof.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_alt/libbvargpu_prof.so \
     $(ls build/*.o | grep -v bvg_kern_tvp_g32.o) ../../build_alt/bvg_kern_tvp_g32_prof.o -lcudart
