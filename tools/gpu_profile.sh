#!/bin/bash
# ncu evidence for one round: launch lists (train step, synthesis) + full captures of the GEMM kernels.
# Usage (under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_train_launches.csv python tools/profile_step.py train bf16 24 > gpurun_out/${TAG}_train.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_synth_launches.csv python tools/profile_step.py synth bf16 4 > gpurun_out/${TAG}_synth.log 2>&1
# forward GEMMs (gate, res, skip) of the synthesis path
$NCU --set full --import-source on -k regex:tc_gemm -s 10 -c 8 -f -o gpurun_out/${TAG}_fwd python tools/profile_step.py synth bf16 4 > gpurun_out/${TAG}_fwd.log 2>&1
# backward GEMMs of the training step: skip the 192 forward + 16 recompute GEMM launches of the last flow
$NCU --set full --import-source on -k regex:tc_ -s 211 -c 8 -f -o gpurun_out/${TAG}_bwd python tools/profile_step.py train bf16 24 > gpurun_out/${TAG}_bwd.log 2>&1
ls -la gpurun_out
# forward GEMMs of the TRAINING step (B=24: the launch bench.py's roofline quotes; gives its DRAM traffic)
$NCU --set full --import-source on -k regex:tc_gemm -s 10 -c 6 -f -o gpurun_out/${TAG}_trainfwd python tools/profile_step.py train bf16 24 > gpurun_out/${TAG}_trainfwd.log 2>&1
ls -la gpurun_out
