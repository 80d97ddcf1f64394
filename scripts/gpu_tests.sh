#!/bin/bash
# GPU test suite (optionally -k "$1"), output to gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -${2:-30} gpurun_out/pytest_gpu.log
