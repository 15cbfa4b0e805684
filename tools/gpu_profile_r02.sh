#!/bin/bash
# ncu evidence for round 2 (run under gpurun): launch lists of one training step / one synthesis call, full captures of the
# three task kernels (forward without saves, forward with saves, backward chain), of the batched weight-gradient kernel and
# of the HBM-bound elementwise kernels.  The .ncu-rep files are summarised ON THE BOX (raw page -> tools/ncu_raw_summary.py,
# source page -> tools/ncu_lines.py) and removed: gpurun copies back at most 64 MiB.
# Usage: bash tools/gpu_profile_r02.sh <tag> [precision]
TAG=${1:-r02}
PREC=${2:-fp16}
O=gpurun_out
mkdir -p $O
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_train_launches.csv python tools/profile_step.py train $PREC 24 > $O/${TAG}_train.log 2>&1
python tools/summarize_launches.py $O/${TAG}_train_launches.csv > $O/${TAG}_train_step_launches.txt
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_synth_launches.csv python tools/profile_step.py synth $PREC 4 > $O/${TAG}_synth.log 2>&1
python tools/summarize_launches.py $O/${TAG}_synth_launches.csv > $O/${TAG}_synth_b4_launches.txt
rm -f $O/${TAG}_train_launches.csv $O/${TAG}_synth_launches.csv
summ() {  # $1 = report stem
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py $O/$1_raw.csv > $O/$1_summary.txt 2>&1
  if [ "$2" = "src" ]; then
    ncu -i $O/$1.ncu-rep --page source --csv --print-source sass,cuda > $O/$1_source.csv 2>/dev/null
    python tools/ncu_lines.py $O/$1_source.csv 40 >> $O/$1_summary.txt 2>&1
    rm -f $O/$1_source.csv
  fi
  rm -f $O/$1.ncu-rep
}
# task kernels: launch 12 of the step is the last forward without saves, 13 the first recompute (with saves), 14 the first backward chain
$NCU --set full --import-source on -k regex:mega -s 11 -c 3 -f -o $O/${TAG}_mega python tools/profile_step.py train $PREC 24 > $O/${TAG}_mega.log 2>&1
summ ${TAG}_mega src
$NCU --set full -k regex:tc_wgrad -c 2 -f -o $O/${TAG}_wgrad python tools/profile_step.py train $PREC 24 > $O/${TAG}_wgrad.log 2>&1
summ ${TAG}_wgrad
# HBM-bound kernels: achieved DRAM throughput
$NCU --set full -k 'regex:coupling|conv1x1_apply|squeeze|end_fwd|end_bwd|start_bwd|smallk|sum_per_batch|cond_|nll|upsample|flow_|amax|weight_norm|wgrad_reduce' -c 60 -f -o $O/${TAG}_elementwise python tools/profile_step.py train $PREC 24 > $O/${TAG}_elementwise.log 2>&1
summ ${TAG}_elementwise
du -sh $O; ls -la $O | tail -15
