#!/bin/bash
# round-2 session-2 check C: start conv folded into layer 0 of the non-saving forward (run under gpurun)
O=gpurun_out; T=${1:-r03_c}
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; tail -15 $O/${T}_pytest.log
python tools/dbg/mega_time.py > $O/${T}_megatime.log 2>&1; grep -E "median|epilogue R|epilogue G|mma" $O/${T}_megatime.log
CMWG_FOLD0=0 python tools/dbg/mega_time.py > $O/${T}_megatime_fold0.log 2>&1; grep -E "median" $O/${T}_megatime_fold0.log
python bench.py --no-wsrglow --no-waveflow --no-synth-sweep --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; T=$T python - <<'PY'
import json,os
d=json.loads(open("gpurun_out/%s_bench.json" % os.environ["T"]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["kernel_classes_ms_per_step"])
print(d["synth"]["value"], d["synth"]["roofline"], d["clocks"])
PY
