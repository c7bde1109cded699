#!/bin/bash
# Builds libvpt_b200.so (the product) in-tree for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/_obj"
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVCC="${NVCC:-nvcc}"
# kernels: the reference's numeric flags (--use_fast_math) so per-seed parity is reachable
$NVCC $ARCH -O3 --use_fast_math -lineinfo -std=c++17 -Xcompiler -fPIC ${VPT_KERNEL_DEFS:-} -c "$HERE/device/vpt_kernels.cu" -o "$HERE/_obj/vpt_kernels.o"
$NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c "$HERE/device/vpt_bricks.cu" -o "$HERE/_obj/vpt_bricks.o"
# octree build: plain IEEE flags, as the reference's bvh object
$NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -c "$HERE/device/vpt_octree.cu" -o "$HERE/_obj/vpt_octree.o"
$NVCC $ARCH -O2 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off ${VPT_KERNEL_DEFS:-} -x cu -c "$HERE/host/vpt_context.cpp" -o "$HERE/_obj/vpt_context.o"
$NVCC $ARCH -O2 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -x cu -c "$HERE/host/vpt_scene_build.cpp" -o "$HERE/_obj/vpt_scene_build.o"
g++ -O2 -std=c++17 -fPIC -c "$HERE/host/vdb_reader.cpp" -o "$HERE/_obj/vdb_reader.o"
g++ -O2 -std=c++17 -fPIC -c "$HERE/host/image_io.cpp" -o "$HERE/_obj/image_io.o"
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c "$HERE/host/sky_table.cpp" -o "$HERE/_obj/sky_table.o"
LIBNAME="${VPT_LIB_NAME:-libvpt_b200.so}"
$NVCC $ARCH -O2 -std=c++17 -Xcompiler -fPIC -x cu -c "$HERE/host/vpt_comm.cpp" -o "$HERE/_obj/vpt_comm.o"
$NVCC $ARCH -shared -o "$OUT/$LIBNAME" "$HERE"/_obj/*.o -lz -ldl
echo "built $OUT/$LIBNAME"
