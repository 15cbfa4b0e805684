#!/bin/bash
# warm-cache launch list of one training step incl. optimizer + repack (ncu, serialised): bash tools/gpu_launchlist.sh <tag>
O=gpurun_out; T=${1:-r03_ll}
mkdir -p $O
ncu --profile-from-start off --clock-control none --cache-control none --metrics gpu__time_duration.sum --csv --log-file $O/${T}_launches.csv python tools/profile_step.py trainopt fp16 24 > $O/${T}_train.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv > $O/${T}_train_step_launches_warm.txt
rm -f $O/${T}_launches.csv
cat $O/${T}_train_step_launches_warm.txt
