#!/bin/bash
# usage: tools/gpu_ab2.sh TAG "bench args" lib1.so lib2.so ... — one bench line per library build (A/B of variants)
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; ARGS=$2; shift 2
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"), d["build"]["invert_ms"], d["build"]["tile_layout_s"])'
for lib in "$@"; do
  echo -n "$lib [$ARGS]: "; DI_B200_LIB=$lib timeout 400 python bench.py $ARGS 2>>$O/ab_$TAG.err | tee -a $O/ab_$TAG.jsonl | python -c "$P"
done 2>&1 | tee $O/ab_$TAG.txt
echo total $SECONDS s
