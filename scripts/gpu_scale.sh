#!/bin/bash
# weak-scaling pass on an N-GPU box: bench.py at 4 and 8 ranks (and 2), config 5 (sharded CG) at 8 ranks
mkdir -p gpurun_out
for N in ${1:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench_configs.py 5 119 2 100 > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err; echo "config5 N=$N rc=$?"
  cat gpurun_out/config5_n$N.json
done
