#!/bin/bash
# ncu --set full of the slow small kernels of the backward, with per-line stall samples (run under gpurun)
O=gpurun_out; T=${1:-r03_small}
mkdir -p $O
ncu --profile-from-start off --clock-control none --set full --import-source on -k 'regex:start_bwd256|end_bwd_dw256|weight_norm_bwd|smallk_to_slab|cond_unpack|pack_foldend|weight_eff' -c 9 -f -o $O/${T} python tools/profile_step.py trainopt fp16 24 > $O/${T}.log 2>&1
ncu -i $O/${T}.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${T}_raw.csv > $O/${T}_summary.txt 2>&1
ncu -i $O/${T}.ncu-rep --page source --csv --print-source sass,cuda > $O/${T}_source.csv 2>/dev/null
python tools/ncu_lines.py $O/${T}_source.csv 60 >> $O/${T}_summary.txt 2>&1
python - <<'PY' >> $O/${T}_summary.txt
import csv,sys,os
T=os.environ.get("T","r03_small")
rows=list(csv.reader(open("gpurun_out/%s_raw.csv"%T)))
hdr=rows[0]
want=[i for i,h in enumerate(hdr) if any(k in h for k in ("Kernel Name","achieved_occupancy","warps_active.avg.pct","issue_active.avg.pct","stalled_long_scoreboard","stalled_barrier","stalled_short_scoreboard","stalled_lg_throttle","stalled_math_pipe","stalled_wait.","stalled_mio_throttle","inst_executed.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","registers_per_thread","shared_mem_per_block"))]
for r in rows[2:]:
    print([ (hdr[i][:60], r[i][:40]) for i in want])
PY
rm -f $O/${T}_source.csv $O/${T}.ncu-rep
tail -100 $O/${T}_summary.txt
