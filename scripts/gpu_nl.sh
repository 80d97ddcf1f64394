#!/bin/bash
# nonlinear path: parity tests, then full-size configs 3 and 4 (ms per Newton assembly)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nonlinear or neohooke or rcd or stvenant or example" > gpurun_out/pytest_nl.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_nl.log
for c in 3 4; do timeout 600 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/config$c.json')); print({k:d[k] for k in ('config','ms','phase_ms','checks')}, {k:d[k] for k in d if k in ('roofline','parity')})"; tail -3 gpurun_out/config$c.err; done
