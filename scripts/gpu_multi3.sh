#!/bin/bash
# short 2-GPU check of a build: NCCL parity tests of the sharded paths + the weak- and strong-scaling bench lines (strong with the
# entrywise comparison against the one-GPU matrix and rhs)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -s > gpurun_out/pytest_dist_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_dist_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711"
for sc in weak strong; do
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/bench_${sc}_n$N.json 2> gpurun_out/bench_${sc}_n$N.err; echo "bench $sc rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${sc}_n$N.json").read().strip().splitlines()[-1])
    print("$sc", "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gcells/s", round(d["value"] / 1e9, 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), d.get("parity_vs_1gpu"), d["phase_ms"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_${sc}_n$N.err").read()[-2000:])
PY
done
