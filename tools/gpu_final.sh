#!/bin/bash
# usage: tools/gpu_final.sh TAG — the round-end sequence on one GPU: tests, smoke(), the default bench line (as the driver runs
# it), the reference arm, configs[3] on one GPU, then the profiling pass.
set -u
O=gpurun_out; mkdir -p $O
TAG=$1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err; echo "reference arm rc=$?"
timeout 900 python bench.py --workload c4 --steps 3 --cpu-sample 32 > $O/bench_${TAG}_c4_n1.json 2> $O/bench_${TAG}_c4_n1.err; echo "c4 rc=$?"
python - <<PY
import json
for n in ("n1","ref","c4_n1"):
    try:
        d=json.loads(open("$O/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d.get("ms_per_step"), d["e2e"]["value"], (d.get("parity") or {}).get("bit_exact"), (d.get("api_e2e") or {}).get("vs_e2e"), d.get("build"))
    except Exception as e:
        print(n, "FAILED", e)
PY
[ -n "${DI_SKIP_PROF:-}" ] || bash tools/gpu_prof.sh $TAG   # DI_SKIP_PROF=1: the captures were taken by an earlier call
echo total $SECONDS s
