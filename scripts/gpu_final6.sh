#!/bin/bash
# state-independent Jacobian entries no longer travel between the point kernel and the contraction kernel: full GPU suite,
# configs 3 and 4 (full-size parity), config 3 with the option off in the same build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for c in 3 4; do
  timeout 200 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"
  python -c "
import json; d = json.load(open('gpurun_out/config$c.json')); print('config', $c, 'ms', round(d['ms'], 3), d['phase_ms'], d.get('parity'))"
done
EXTFEM_NO_PARITY=1 EXTFEM_OPTIONS="nonlinear_const_jacobian=0" timeout 200 python bench_configs.py 3 > gpurun_out/config3_old.json 2> gpurun_out/config3_old.err; echo "config 3 (option off) rc=$?"
python -c "
import json; d = json.load(open('gpurun_out/config3_old.json')); print('config 3 option off: ms', round(d['ms'], 3), d['phase_ms'])"
