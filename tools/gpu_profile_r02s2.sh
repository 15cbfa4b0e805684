#!/bin/bash
# ncu evidence for round 2, second session (run under gpurun): cold launch list of one training step (the standard
# --metrics gpu__time_duration.sum --clock-control none pass), full captures of the three task kernels (with per-line stall
# samples) and of the two weight-gradient kernels.  Reports are summarised on the box and removed (64 MiB copy-back limit).
TAG=${1:-r02_s2}
PREC=${2:-fp16}
O=gpurun_out
mkdir -p $O
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_train_launches.csv python tools/profile_step.py trainopt $PREC 24 > $O/${TAG}_train.log 2>&1
python tools/summarize_launches.py $O/${TAG}_train_launches.csv > $O/${TAG}_train_step_launches.txt
rm -f $O/${TAG}_train_launches.csv
summ() {  # $1 = report stem
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py $O/$1_raw.csv > $O/$1_summary.txt 2>&1
  if [ "$2" = "src" ]; then
    ncu -i $O/$1.ncu-rep --page source --csv --print-source sass,cuda > $O/$1_source.csv 2>/dev/null
    python tools/ncu_lines.py $O/$1_source.csv 30 >> $O/$1_summary.txt 2>&1
    rm -f $O/$1_source.csv
  fi
  rm -f $O/$1.ncu-rep $O/$1_raw.csv
}
# task kernels: launch 12 of the step is the last forward without saves, 13 the first recompute (with saves), 14 the first backward chain
$NCU --set full --import-source on -k regex:mega -s 11 -c 3 -f -o $O/${TAG}_mega python tools/profile_step.py train $PREC 24 > $O/${TAG}_mega.log 2>&1
summ ${TAG}_mega src
$NCU --set full -k regex:tc_wgrad -c 2 -f -o $O/${TAG}_wgrad python tools/profile_step.py train $PREC 24 > $O/${TAG}_wgrad.log 2>&1
summ ${TAG}_wgrad
head -8 $O/${TAG}_train_step_launches.txt; head -6 $O/${TAG}_mega_summary.txt; cat $O/${TAG}_wgrad_summary.txt | head -5
