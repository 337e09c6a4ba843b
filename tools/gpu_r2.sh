#!/bin/bash
# usage: tools/gpu_r2.sh TAG [lib.so ...] — GPU tests, then bench lines (both workloads) for the in-tree library
# and every extra library given (A/B of builds of the same sources)
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; shift
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"), d["build"]["invert_ms"], d["build"]["tile_layout_s"], d["index"]["payload_gb"], d["index"]["dense_posting_frac"], d["postings_per_query"])'
t0=$SECONDS
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -5 $O/pytest_$TAG.log
fi
run() { echo -n "$1 [$2]: "; DI_B200_LIB=$1 timeout 400 python bench.py $2 2>>$O/sweep_$TAG.err | tee -a $O/bench_$TAG.jsonl | python -c "$P"; }
{
for lib in improving-learned-index_b200/libdi_b200.so "$@"; do
run $lib "--steps 5 --cpu-sample 32"
run $lib "--steps 5 --cpu-sample 0 --unique-terms 0"
done
run improving-learned-index_b200/libdi_b200.so "--workload c4 --steps 3 --cpu-sample 32"
} 2>&1 | tee $O/sweep_$TAG.txt
echo total $SECONDS s
