#!/bin/bash
# usage: tools/gpu_multi.sh TAG "N1 N2 ..." [extra bench args] — peer-exchange tests on all visible GPUs, then for every N
# the bench with the fused peer-memory exchange and with the NCCL all-gather fallback (DI_B200_NO_PEER=1)
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; NS=$2; shift 2
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], d.get("exchange","")[:20], d.get("second_pass_queries_rank0_last_step"), (d.get("sharded_parity") or {}).get("bit_exact"))'
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/pytest_peer_$TAG.log 2>&1; echo "peer tests rc=$?"; tail -4 $O/pytest_peer_$TAG.log
port=29500
for n in $NS; do
  for mode in fused nccl; do
    port=$((port+1))
    if [ $mode = nccl ]; then export DI_B200_NO_PEER=1; else unset DI_B200_NO_PEER; fi
    echo -n "N=$n $mode: "
    if [ $n = 1 ]; then
      [ $mode = nccl ] && { echo skip; continue; }
      timeout 600 python bench.py --gpus 1 --steps 10 --cpu-sample 0 --py-ref-seconds 0 "$@" 2>>$O/multi_$TAG.err | tee -a $O/multi_$TAG.jsonl | python -c "$P"
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --cpu-sample 0 --py-ref-seconds 0 --verify-sharded "$@" 2>>$O/multi_$TAG.err | tee -a $O/multi_$TAG.jsonl | python -c "$P"
    fi
  done
done 2>&1 | tee $O/multi_$TAG.txt
echo total $SECONDS s
