#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_train_launches.csv python tools/profile_step.py train bf16 24 > gpurun_out/${TAG}_train.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_synth_launches.csv python tools/profile_step.py synth bf16 4 > gpurun_out/${TAG}_synth.log 2>&1
