#!/bin/bash
# ncu --set full of the slow small kernels of the backward: one compact line per kernel + the top stall lines (run under gpurun)
O=gpurun_out; T=${1:-r03_small}; PAT=${2:-'start_bwd256|foldend_dw|wgrad_reduce|weight_norm_bwd'}; N=${3:-4}
mkdir -p $O
ncu --profile-from-start off --clock-control none --set full --import-source on -k "regex:$PAT" -c $N -f -o $O/${T} python tools/profile_step.py trainopt fp16 24 > $O/${T}.log 2>&1
ncu -i $O/${T}.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
T=$T python - <<'PY' > $O/${T}_summary.txt
import csv,os
T=os.environ["T"]
rows=list(csv.reader(open("gpurun_out/%s_raw.csv"%T)))
hdr=rows[0]
def col(name):
    for i,h in enumerate(hdr):
        if h==name: return i
    for i,h in enumerate(hdr):
        if name in h: return i
    return None
keys=[("Kernel Name","name"),("gpu__time_duration.sum","ns"),("smsp__inst_executed.sum","inst"),("sm__issue_active.avg.pct_of_peak_sustained_elapsed","issue%"),
("sm__warps_active.avg.pct_of_peak_sustained_active","occ%"),("dram__bytes_read.sum","rdB"),("dram__bytes_write.sum","wrB"),("lts__t_bytes.sum","l2B"),
("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","long_sb"),("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","short_sb"),
("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","barrier"),("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","mio"),
("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","lg"),("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","wait"),
("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","math"),("launch__registers_per_thread","regs"),("launch__grid_size","grid")]
idx=[(col(k),n) for k,n in keys]
for r in rows[2:]:
    print("  ".join("%s=%s"%(n,(r[i][:46] if i is not None else "?")) for i,n in idx))
PY
ncu -i $O/${T}.ncu-rep --page source --csv --print-source sass,cuda > $O/${T}_source.csv 2>/dev/null
python tools/ncu_lines.py $O/${T}_source.csv 28 >> $O/${T}_summary.txt 2>&1
rm -f $O/${T}_source.csv $O/${T}.ncu-rep $O/${T}_raw.csv
cat $O/${T}_summary.txt
