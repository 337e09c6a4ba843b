#!/bin/bash
# GPU tests of the in-tree library, then an A/B of library variants (one bench line each)
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; shift
t0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -5 $O/pytest_$TAG.log
bash tools/gpu_ab2.sh $TAG '--steps 5 --cpu-sample 16 --no-file-legs' "$@"
