#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --set full --import-source on -k regex:tc_gemm -s 10 -c 4 -f -o gpurun_out/${TAG}_fwd python tools/profile_step.py synth bf16 4 > gpurun_out/${TAG}_fwd.log 2>&1
tail -3 gpurun_out/${TAG}_fwd.log
