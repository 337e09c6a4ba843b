#!/bin/bash
# usage: tools/gpu_full.sh TAG [variant libs...] — GPU tests, the default bench line with every leg, the reference arm, then an A/B
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; shift
t0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -5 $O/pytest_$TAG.log
t0=$SECONDS
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$? $((SECONDS-t0))s"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["score_ms_per_step"], d.get("parity"), d["build"])
print(d.get("cpu_baseline")); print(d.get("cpu_baseline_python")); print(d.get("api_e2e")); print(d.get("file_legs_note"))
PY
if [ $# -gt 0 ]; then bash tools/gpu_ab2.sh $TAG '--steps 5 --cpu-sample 16 --no-file-legs' "$@"; fi
echo total $SECONDS s
