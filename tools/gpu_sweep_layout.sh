P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["roofline"]["score_ms_per_step"], d["index"]["payload_gb"], d["index"]["dense_posting_frac"], d["parity"]["bit_exact"])'
for dr in 4 6 8 10 12 16; do echo -n "dense_ratio $dr: "; python bench.py --steps 5 --cpu-sample 16 --no-file-legs --dense-ratio $dr 2>/dev/null | python -c "$P"; done
for cs in 1500 3000; do echo -n "cand_slack $cs: "; python bench.py --steps 5 --cpu-sample 16 --no-file-legs --cand-slack $cs 2>/dev/null | python -c "$P"; done
