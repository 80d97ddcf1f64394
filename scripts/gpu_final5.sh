#!/bin/bash
# point kernel of the nonlinear path after the shared-memory cache of the physical basis values and the row-wise Neo-Hooke
# Jacobian: full GPU suite, configs 4 and 3 (with the full-size parity check), config 4 with both options off in the same build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for c in 4 3; do
  timeout 300 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"
  python -c "
import json; d = json.load(open('gpurun_out/config$c.json')); print('config', $c, 'ms', round(d['ms'], 3), d['phase_ms'], d.get('parity'), d.get('kernel_ms'))"
done
EXTFEM_NO_PARITY=1 EXTFEM_OPTIONS="nonlinear_point_cache=0,nonlinear_rowwise=0" timeout 300 python bench_configs.py 4 > gpurun_out/config4_old.json 2> gpurun_out/config4_old.err; echo "config 4 (options off) rc=$?"
python -c "
import json; d = json.load(open('gpurun_out/config4_old.json')); print('config 4 options off: ms', round(d['ms'], 3), d['phase_ms'], d.get('kernel_ms'))"
