#!/bin/bash
# Developer A/B builds of the trace kernel: tools/build_variant.sh TAG "-DRG_X=1 ..."  ->  build/variants/librgb200_TAG.so
# (select it with RGB200_LIB=build/variants/librgb200_TAG.so python tools/gpu_time.py c3)
# WITH_BUILD=1: the defines also go to rg_build.cu (layout switches of rg_types.cuh such as RG_HALF_SLAB)
set -e
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p build/variants
C=raygun_b200/csrc
make -C $C -s >/dev/null 2>&1
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -fmad=false -Xptxas -v "$@" -c ${SRC:-$C/rg_trace.cu} -o build/variants/rg_trace_$TAG.o 2>&1 | grep -A2 "k_trace_\(pool\|lanes\)ILb0ELb0" | grep -E "spill|Used" || true
BUILD_O=$C/_obj/rg_build.o
if [ -n "$WITH_BUILD" ]; then
  /usr/local/cuda/bin/nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $C/rg_build.cu -o build/variants/rg_build_$TAG.o
  BUILD_O=build/variants/rg_build_$TAG.o
fi
/usr/local/cuda/bin/nvcc $ARCH -shared -o build/variants/librgb200_$TAG.so $C/_obj/rg_api.o $BUILD_O $C/_obj/rg_post.o $C/_obj/rg_scene.o build/variants/rg_trace_$TAG.o -lcudart
