#!/bin/bash
# multi-GPU pass ($1 = number of GPUs): NCCL parity test on 2 ranks, then the bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -s > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"
tail -6 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
