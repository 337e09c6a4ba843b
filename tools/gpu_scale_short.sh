set -u
O=gpurun_out
P="import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['score_ms_per_step'], d['roofline']['finalize_ms_per_step'], (d.get('sharded_parity') or {}).get('bit_exact'))"
port=29720
for n in 8 4; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --verify-sharded --cpu-sample 16 --no-file-legs --py-ref-seconds 0 > $O/bench_r7_n$n.json 2> $O/bench_r7_n$n.err; echo "N=$n rc=$?"; python -c "$P" $O/bench_r7_n$n.json
done
port=$((port+1))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port tools/sharded_timeline.py 2>>$O/scale_r7.err | tee $O/timeline_r7_n8.txt | tail -4
echo total $SECONDS s
