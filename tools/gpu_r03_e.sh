#!/bin/bash
# round-2 session-2 check E: vector pack, single weight-norm backward launch; parity + warm launch list (run under gpurun)
O=gpurun_out; T=${1:-r03_e}
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
python bench.py --no-wsrglow --no-waveflow --no-synth-sweep > $O/${T}_bench.json 2> $O/${T}_bench.err; T=$T python - <<'PY'
import json,os
d=json.loads(open("gpurun_out/%s_bench.json" % os.environ["T"]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["kernel_classes_ms_per_step"])
print(d.get("parity")); print(d["synth"]["value"], d["clocks"])
PY
bash tools/gpu_launchlist.sh ${T}_ll
