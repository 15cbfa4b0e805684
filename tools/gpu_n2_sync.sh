#!/bin/bash
# N = 2: gradient exchange overlapped with the backward (one all-reduce per flow) vs one all-reduce at its end.
# Every launch is bounded by its own timeout (a hung rank must not hold the box).
O=gpurun_out; T=${1:-r03_n2}; MODES=${2:-"deferred overlap"}
mkdir -p $O
for mode in $MODES; do
  CMWG_GRAD_SYNC=$mode timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-synth --no-wsrglow --no-waveflow --no-cpu-baseline --no-strong > $O/${T}_${mode}.json 2> $O/${T}_${mode}.err
  echo "$mode exit $?"
  M=$mode T=$T python - <<'PY'
import json,os
p="gpurun_out/%s_%s.json" % (os.environ["T"], os.environ["M"])
try:
    d=json.loads([l for l in open(p) if l.startswith("{")][-1])
    print(os.environ["M"], {k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["kernel_classes_ms_per_step"], d["clocks"])
except Exception as e:
    print(os.environ["M"], "failed", e); print(open(p.replace(".json",".err")).read()[-1500:])
PY
done
