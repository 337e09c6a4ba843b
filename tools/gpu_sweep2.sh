#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
D=improving-learned-index_b200
TAG=$1
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], (d.get("parity") or {}).get("bit_exact"))'
run() { echo -n "$1 [$2]: "; DI_B200_LIB=$1 timeout 300 python bench.py --steps 5 --cpu-sample 0 $2 2>>$O/sweep_$TAG.err | python -c "$P"; }
{
run $D/libdi_b200.so ""
run $D/variants/libdi_t256_b3.so "--tile-docs 32768"
run $D/variants/libdi_su8.so ""
run $D/variants/libdi_su2.so ""
run $D/variants/libdi_b5.so ""
run $D/libdi_b200.so "--tile-docs 32768"
run $D/libdi_b200.so "--tile-docs 8192"
} 2>&1 | tee $O/sweep_$TAG.txt
