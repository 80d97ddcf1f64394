#!/bin/bash
# N-GPU record: [NCCL parity tests of the sharded paths,] bench weak + strong, config 5.  usage: gpu_multi2.sh N [n] [notest]
N=${1:-2}; n=${2:-119}
mkdir -p gpurun_out
if [ -z "$3" ]; then
  timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -s > gpurun_out/pytest_dist_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dist_n$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711"
for sc in weak strong; do
  timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --grid-n $n --scaling $sc > gpurun_out/bench_${sc}_n$N.json 2> gpurun_out/bench_${sc}_n$N.err; echo "bench $sc rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${sc}_n$N.json").read().strip().splitlines()[-1])
    print("$sc", "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gcells/s", round(d["value"] / 1e9, 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), d.get("parity_vs_1gpu"), d.get("local"), d["phase_ms"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_${sc}_n$N.err").read()[-2000:])
PY
done
timeout 900 $TR bench_configs.py 5 $n > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err; echo "config5 rc=$?"; tail -c 900 gpurun_out/config5_n$N.json; tail -2 gpurun_out/config5_n$N.err
