#!/bin/bash
# One GPU call of the build→measure loop: parity tests, the bench line, an A/B sweep of library variants,
# a shard-sized run, and the per-phase cycle profile (diagnostic build). Outputs under gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
D=improving-learned-index_b200
TAG=${1:-r1c}
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"])'
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -3 $O/pytest_$TAG.log
t0=$SECONDS
timeout 600 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err; echo "bench rc=$? $((SECONDS-t0))s"; cat $O/bench_${TAG}_n1.json | python -c "$P"
run() { echo -n "$1 [$2]: "; DI_B200_LIB=$1 timeout 300 python bench.py --steps 3 --cpu-sample ${3:-0} $2 2>>$O/sweep_$TAG.err | python -c "$P"; }
{
for lib in $D/libdi_b200.so $D/variants/libdi_head.so $D/variants/libdi_nodefer.so; do
  run $lib "" 16
  run $lib "--docs 1105228"
done
run $D/libdi_b200.so "--docs 1105228 --top-k 221"
run $D/variants/libdi_head.so "--docs 1105228 --top-k 221"
} 2>&1 | tee $O/sweep_$TAG.txt
DI_B200_PROF=$O/phases_${TAG}_full.csv run $D/variants/libdi_prof.so "--steps 1 --warmup 3"
DI_B200_PROF=$O/phases_${TAG}_shard.csv run $D/variants/libdi_prof.so "--steps 1 --warmup 3 --docs 1105228"
echo total $SECONDS s
