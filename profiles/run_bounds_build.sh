#!/bin/bash
# Memory-safety evidence without compute-sanitizer (closed on this GPU pool): build libtbirch with
# -DTB_DEBUG_BOUNDS (device-side asserts on every computed shared-memory index / output position of the
# row kernels and on the lookup's table index) and run the GPU parity tests against that build.
#   here (no GPU):   profiles/run_bounds_build.sh build
#   on the GPU box:  profiles/run_bounds_build.sh test <tag>
set -e
OUT=ternary_birch_b200/libtbirch_bounds.so
if [ "$1" = "build" ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DTB_DEBUG_BOUNDS \
       ternary_birch_b200/csrc/tbirch.cu -o $OUT
  ls -la $OUT
else
  TAG=${2:-r2}
  TBIRCH_LIB=$PWD/$OUT python -m pytest tests/test_gpu_parity.py tests/test_genus_build.py -x -q -m gpu \
      > gpurun_out/${TAG}_bounds_build_pytest.log 2>&1 || true
  tail -3 gpurun_out/${TAG}_bounds_build_pytest.log
fi
