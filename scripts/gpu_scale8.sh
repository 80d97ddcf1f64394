mkdir -p gpurun_out
for N in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
