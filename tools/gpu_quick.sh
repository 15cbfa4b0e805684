#!/bin/bash
# Staged GPU check: engine self test first (short timeout), the rest only if it passes.
mkdir -p gpurun_out
timeout 150 python -W ignore -m pytest tests/test_gpu_tc_engine.py -x -q -m gpu --timeout=60 > gpurun_out/test_gpu_tc_engine.log 2>&1
rc=$?
tail -n 30 gpurun_out/test_gpu_tc_engine.log
if [ $rc -ne 0 ]; then echo "tc engine self test failed (rc=$rc); skipping the rest"; exit 1; fi
for f in tests/test_gpu_primitives.py tests/test_gpu_wn.py tests/test_gpu_modules.py; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout 400 python -W ignore -m pytest $f -q -m gpu --maxfail=8 --timeout=120 > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -n 25 gpurun_out/$name.log
done
if [ "$1" == "bench" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  tail -c 2500 gpurun_out/bench_q.json; tail -3 gpurun_out/bench_q.err
fi
