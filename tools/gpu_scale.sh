#!/bin/bash
# usage (on an 8-GPU box): tools/gpu_scale.sh TAG — the scaling run the driver does (N = 1, 2, 4, 8, same command lines), the
# per-stage timeline of a sharded step for N = 2, 4, 8, configs[3] on 8 GPUs, and the NCCL-fallback form at N = 8.
set -u
O=gpurun_out; mkdir -p $O
TAG=$1
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["score_ms_per_step"], d["roofline"]["finalize_ms_per_step"], str(d.get("exchange",""))[:12], d.get("second_pass_queries_rank0_last_step"), (d.get("sharded_parity") or {}).get("bit_exact"), (d.get("parity") or {}).get("bit_exact"))'
port=29600
run() { # n, extra args, tag
  port=$((port+1))
  if [ $1 = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 $2 2>>$O/scale_$TAG.err | tee $O/bench_${TAG}_$3.json | python -c "$P"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $port bench.py --gpus $1 --steps 10 --warmup 3 $2 2>>$O/scale_$TAG.err | tee $O/bench_${TAG}_$3.json | python -c "$P"
  fi
}
{
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -2
echo -n "N=2: "; run 2 "" n2
for n in 4 8; do echo -n "N=$n: "; run $n "--verify-sharded" n$n; done
echo -n "N=8 nccl fallback: "; DI_B200_NO_PEER=1 run 8 "" n8_nccl
echo -n "N=8 c4: "; run 8 "--workload c4 --steps 3" c4_n8
for n in 2 4 8; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port tools/sharded_timeline.py 2>>$O/scale_$TAG.err | tee $O/timeline_${TAG}_n$n.txt
done
} 2>&1 | tee $O/scale_$TAG.txt
echo total $SECONDS s
