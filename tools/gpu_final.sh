#!/bin/bash
# final validation of the session on one box: GPU test suite, smoke, both bench arms (run under gpurun)
O=gpurun_out; T=${1:-r02_s2_final}
mkdir -p $O
timeout 300 python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; tail -2 $O/${T}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -3 $O/${T}_smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; cut -c1-300 $O/${T}_bench_ref.json | tail -1
timeout 420 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; T=$T python - <<'PY'
import json,os
d=json.loads([l for l in open("gpurun_out/%s_bench.json" % os.environ["T"]) if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["kernel_classes_ms_per_step"])
print("parity", {k:v for k,v in d.get("parity",{}).items() if k.endswith("l2") or k.startswith("grad")})
s=d["synth"]; print("synth", s["value"], s["e2e"]["value"], s["roofline"]["frac"], s.get("per_gpu_khz_by_batch"), s.get("x_cpu_per_gpu"))
w=d.get("wsrglow",{}); print("wsrglow", w.get("train_segments_per_s"), w.get("inverse_khz"), w.get("train_frac_of_peak"), w.get("gpu_launches_per_step"))
f=d.get("waveflow",{}); print("waveflow", f.get("train_segments_per_s"), f.get("synth_khz"), f.get("synth_ms"))
print("clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
