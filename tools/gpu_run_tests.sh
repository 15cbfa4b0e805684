#!/bin/bash
# Run every GPU test file in its own process with a timeout so that one hang or sticky CUDA error
# cannot take the others down; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
for f in tests/test_gpu_tc_engine.py tests/test_gpu_primitives.py tests/test_gpu_wn.py tests/test_gpu_modules.py; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout ${TEST_TIMEOUT:-600} python -W ignore -m pytest $f -q -m gpu --maxfail=${MAXFAIL:-8} --timeout=300 ${PYTEST_EXTRA} > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -n 25 gpurun_out/$name.log
done
