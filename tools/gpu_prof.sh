#!/bin/bash
# usage: tools/gpu_prof.sh TAG — per-phase cycle profile (diagnostic build), launch list and one `ncu --set full`
# capture of the score kernel, each only after the same command exited 0 without ncu
set -u
O=gpurun_out; mkdir -p $O
TAG=$1
D=improving-learned-index_b200
CMD="python bench.py --steps 1 --cpu-sample 0 --py-ref-seconds 0"
DI_B200_PROF=$O/phases_${TAG}_full.csv DI_B200_LIB=$D/variants/libdi_prof.so $CMD > $O/prof_$TAG.log 2>&1; echo "phase profile rc=$?"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_persistent -s 3 -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu2_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 $O/ncu2_$TAG.log
echo total $SECONDS s
