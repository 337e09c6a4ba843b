#!/bin/bash
# Builds the diagnostic library variants used by tools/gpu_prof.sh and tools/gpu_ab2.sh (git-ignored; they travel to the GPU box):
#   variants/libdi_prof.so   -DDI_PROFILE_PHASES: per-tile, per-phase cycle counters ($DI_B200_PROF=file.csv)
# Extra variants: tools/build_variants.sh NAME "-DDI_SCORE_THREADS=256 -DDI_SCORE_MIN_BLOCKS=3" ...
set -eu
cd "$(dirname "$0")/../improving-learned-index_b200"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
mkdir -p variants
nvcc $F -DDI_PROFILE_PHASES -o variants/libdi_prof.so csrc/di_b200.cu csrc/collection.cu csrc/run_io.cu
while [ $# -ge 2 ]; do
  nvcc $F $2 -o variants/libdi_$1.so csrc/di_b200.cu csrc/collection.cu csrc/run_io.cu
  shift 2
done
ls -la variants
