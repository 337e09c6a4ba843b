#!/bin/bash
# usage (on the GPU box): tools/variant_sweep.sh  -> one line per build variant / parameter of the score kernel
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["score_ms_per_step"], d["roofline"]["frac"], d["index"]["payload_gb"], d["index"]["dense_posting_frac"])'
run() { echo -n "$1 $2 $3: "; env $3 DI_B200_LIB=$1 timeout 300 python bench.py --steps 3 --cpu-sample 0 $2 2>/dev/null | python -c "$P"; }
D=improving-learned-index_b200
for cs in 1100 1280 1536 2048; do run $D/libdi_b200.so "--cand-slack $cs" ""; done
for v in $D/variants/*.so; do [ -f $v ] && run $v "" ""; done
