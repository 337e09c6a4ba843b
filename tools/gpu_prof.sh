#!/bin/bash
# usage: tools/gpu_prof.sh TAG — per-phase cycle profile (diagnostic build), launch list, and `ncu --set full` captures of
# the score kernel (one launch = one step), the final select+sort, the radix-sort kernels and the fused merge; each ncu run
# only after the same command exited 0 without ncu. Numbers printed under ncu are never bench values.
set -u
O=gpurun_out; mkdir -p $O
TAG=$1
D=improving-learned-index_b200
CMD="python bench.py --steps 1 --cpu-sample 0 --no-file-legs"
DI_B200_PROF=$O/phases_${TAG}_full.csv DI_B200_LIB=$D/variants/libdi_prof.so $CMD > $O/prof_$TAG.log 2>&1; echo "phase profile rc=$?"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
cap() { # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $O/prof_${TAG}_$1 $CMD > $O/ncu_${TAG}_$1.log 2>&1
  echo "capture $1 rc=$?"
}
cap score score_persistent 3 1
cap finalize finalize_topk 3 1
cap sort 'rs_onesweep|rs_histogram' 0 4
PEER="python -m pytest tests/test_gpu_peer.py -q -m gpu -k pull_merge"
$PEER > $O/plain_${TAG}_peer.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:merge_pull -c 2 -f -o $O/prof_${TAG}_merge $PEER > $O/ncu_${TAG}_merge.log 2>&1
echo "capture merge rc=$?"
echo total $SECONDS s
