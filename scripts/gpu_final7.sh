#!/bin/bash
# last check of the final build: full GPU suite + config 3
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 100 python bench_configs.py 3 > gpurun_out/config3.json 2> gpurun_out/config3.err; echo "config 3 rc=$?"
python -c "
import json; d = json.load(open('gpurun_out/config3.json')); print('config 3 ms', round(d['ms'], 3), d['phase_ms'], d.get('parity'))"
