#!/bin/bash
# compute-sanitizer over the hot path: memcheck + racecheck on smoke() and on the
# row-kernel variants (fused look-back, two-pass, global-scratch) -- the kernels with shared-memory
# atomics, a decoupled look-back spin and a barrier-less sticky flag.  Logs -> gpurun_out/.
TAG=${1:-r2}
export PYTHONUNBUFFERED=1
run() {  # name tool command...
  name=$1; tool=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/${TAG}_sanitizer_${name}_${tool}.log "$@" > gpurun_out/${TAG}_sanitizer_${name}_${tool}.out 2>&1
  echo "$name $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitizer_${name}_${tool}.log | tail -1)"
}
run smoke memcheck python __graft_entry__.py smoke
run smoke racecheck python __graft_entry__.py smoke
run rows memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "row_kernel_variants_agree or row_kernel_global_scratch"
run rows racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "row_kernel_variants_agree or row_kernel_global_scratch"
