#!/bin/bash
# usage: tools/gpu_sweep.sh TAG "bench args" lib1 lib2 ...   (A/B of library builds; one line per run)
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; ARGS=$2; shift 2
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"))'
for lib in "$@"; do
  for a in "" "--docs 1105228" "--docs 1105228 --top-k 221"; do
    echo -n "$lib [$ARGS $a]: "
    DI_B200_LIB=$lib timeout 300 python bench.py --steps 3 $ARGS $a 2>>$O/sweep_$TAG.err | python -c "$P"
  done
done 2>&1 | tee $O/sweep_$TAG.txt
