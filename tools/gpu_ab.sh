#!/bin/bash
# usage: tools/gpu_ab.sh TAG lib1.so lib2.so ...  — parity tests on the default library, then an A/B of builds
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; shift
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"))'
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -3 $O/pytest_$TAG.log
run() { echo -n "$1 [$2]: "; DI_B200_LIB=$1 timeout 300 python bench.py --steps 5 $2 2>>$O/sweep_$TAG.err | python -c "$P"; }
{
for lib in "$@"; do
run $lib "--cpu-sample 16"
run $lib "--cpu-sample 0 --docs 1105228"
run $lib "--cpu-sample 0 --docs 1105228 --top-k 221"
done
} 2>&1 | tee $O/sweep_$TAG.txt
echo total $SECONDS s
