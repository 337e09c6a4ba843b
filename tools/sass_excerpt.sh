#!/bin/bash
# usage: tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt — what the shipped library compiles to (no GPU needed)
set -u
D=improving-learned-index_b200
K='_ZN2di23score_persistent_kernelILb0ELb0EEEvNS_10SearchArgsEPy'
echo "# SASS of $D/libdi_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3), kernel score_persistent_kernel<ACC32=false, BOUNDS=false>"
cuobjdump -sass -fun "$K" $D/libdi_b200.so 2>/dev/null > /tmp/_k3.sass
echo "## instruction histogram (static)"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" /tmp/_k3.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -24
echo
echo "## fused dense + threshold pass, 2 dense segments, predicate-free form: one step of the loop"
echo "##   LDG.E.128.CONSTANT (2 segments x 4 units in flight) -> LDS.128 accumulators -> PRMT (alu pipe) + IMAD.IADD (fma pipe)"
echo "##   -> VIMNMX3.U16x2 / VIMNMX.U16x2 threshold test in registers -> STS.128 of zeros, then a predicated STS.128 of the sums of a hit group"
L=$(grep -n "VIMNMX3" /tmp/_k3.sass | head -1 | cut -d: -f1)
sed -n "$((L-70)),$((L+30))p" /tmp/_k3.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' | cut -c1-100
echo
echo "## sparse phase: LDG.E.128 (L1-allocating) + 4 x ATOMS.ADD per thread"
L=$(grep -n "ATOMS" /tmp/_k3.sass | head -1 | cut -d: -f1)
sed -n "$((L-12)),$((L+14))p" /tmp/_k3.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' | cut -c1-100
echo
echo "## Blackwell/Hopper-class instructions per build (whole library)"
for lib in $D/libdi_b200.so $D/variants/libdi_pf.so $D/variants/libdi_tma.so; do
  [ -f $lib ] || continue
  echo -n "$lib: "
  cuobjdump -sass $lib 2>/dev/null | grep -oE "UBLKPF|UBLKCP|SYNCS[A-Z.]*|MATCH[A-Z.]*|VIMNMX3|REDUX[A-Z.]*" | sort | uniq -c | tr '\n' ' '
  echo
done
echo "# shipped kernel: no TMA instruction — both TMA forms were measured slower (DESIGN.md §4, profiles/r2_k3_ab_record.txt r2j);"
echo "# libdi_pf = -DDI_L2_PREFETCH (UBLKPF = cp.async.bulk.prefetch.L2), libdi_tma = -DDI_DENSE_TMA (UBLKCP = cp.async.bulk.shared::cluster.global, SYNCS = mbarrier)"
