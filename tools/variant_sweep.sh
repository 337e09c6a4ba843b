#!/bin/bash
# usage (on the GPU box): tools/variant_sweep.sh  -> one line per build variant / parameter of the score kernel
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["frac"], d["index"]["payload_gb"], d["index"]["dense_posting_frac"])'
run() { echo -n "$1 $2 $3: "; env $3 DI_B200_LIB=$1 timeout 300 python bench.py --steps 3 --cpu-sample 0 $2 2>/dev/null | python -c "$P"; }
D=improving-learned-index_b200
run $D/variants/libdi_t128_b6.so "--tile-docs 16384" ""
run $D/variants/libdi_t128_b8.so "--tile-docs 8192" ""
run $D/variants/libdi_t128_b8.so "--tile-docs 16384" ""
run $D/variants/libdi_t64_b10.so "--tile-docs 8192" ""
run $D/variants/libdi_t64_b12.so "--tile-docs 8192" ""
run $D/variants/libdi_t64_b12.so "--tile-docs 16384" ""
run $D/variants/libdi_t192_b4.so "--tile-docs 16384" ""
