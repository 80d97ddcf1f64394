#!/bin/bash
# tests matching $1, then bench under the option sets $2.. 
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
shift
bash scripts/gpu_opts.sh "$@"
