#!/bin/bash
# What the driver does at round end, in one call: pytest -m gpu, smoke(), bench (both arms).
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
timeout 1500 python -W ignore -m pytest tests -x -q -m gpu --timeout=300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -n 15 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -W ignore -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
tail -n 5 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -n 5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "ref exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_ref.json; tail -n 5 gpurun_out/${TAG}_bench_ref.err
