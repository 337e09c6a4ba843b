#!/bin/bash
# usage: tools/gpu_ncu.sh TAG  — plain run, then the launch list and one --set full capture of the score kernel
set -u
O=gpurun_out; mkdir -p $O
TAG=$1
CMD="python bench.py --steps 1 --cpu-sample 0"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_persistent -s 3 -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu2_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 $O/ncu2_$TAG.log
