#!/bin/bash
# compute-sanitizer memcheck + initcheck over the GPU parity tests (run through gpurun); summary -> gpurun_out/sanitizer.txt
mkdir -p gpurun_out
out=gpurun_out/sanitizer.txt
sel="visbuffer_bit_exact or alpha or translucent or fixed_exposure_frame_equals"
{
  echo "# compute-sanitizer on the GPU parity tests (B200, gpurun)"
  echo "# command: compute-sanitizer --tool <tool> python -m pytest tests/test_gpu_parity.py tests/test_gpu_fixed_exposure.py -q -x -k '$sel'"
  for tool in memcheck initcheck; do
    echo; echo "## $tool"
    compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_fixed_exposure.py -q -x -k "$sel" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|COMPUTE-SANITIZER|Invalid|Uninitialized" | head -20
  done
} > $out
cat $out
