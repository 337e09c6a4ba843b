#!/bin/bash
# usage: tools/gpu_round2.sh TAG [baseline.so ...]  — tests + A/B of library variants + phase profile
set -u
O=gpurun_out; mkdir -p $O
D=improving-learned-index_b200
TAG=${1:-r1}; shift
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"), d["build"]["tile_layout_s"], d["index"]["payload_gb"], d["index"]["dense_posting_frac"])'
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -3 $O/pytest_$TAG.log
run() { echo -n "$1 $3 [$2]: "; env $3 DI_B200_LIB=$1 timeout 300 python bench.py --steps 3 $2 2>>$O/sweep_$TAG.err | python -c "$P"; }
{
for lib in $D/libdi_b200.so "$@"; do
run $lib "--cpu-sample 16" "X=1"
run $lib "--cpu-sample 0 --docs 1105228" "X=1"
run $lib "--cpu-sample 0 --docs 1105228 --top-k 221" "X=1"
run $lib "--cpu-sample 0 --queries 64" "X=1"
run $lib "--cpu-sample 0 --queries 1" "X=1"
done
} 2>&1 | tee $O/sweep_$TAG.txt
DI_B200_PROF=$O/phases_${TAG}_full.csv run $D/variants/libdi_prof.so "--steps 1 --warmup 3 --cpu-sample 0" "X=1"
DI_B200_PROF=$O/phases_${TAG}_shard.csv run $D/variants/libdi_prof.so "--steps 1 --warmup 3 --cpu-sample 0 --docs 1105228" "X=1"
echo total $SECONDS s
