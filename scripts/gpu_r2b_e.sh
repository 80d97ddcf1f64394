#!/bin/bash
# round 2, second half, run E: full GPU suite + default bench line after the pipelined lower-triangle export
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1.json'))
print({k: round(v, 4) for k, v in d['phase_ms'].items()}, 'step', round(d['ms_per_step'], 4), 'value', d['value'], 'frac', round(d['roofline']['frac'], 4),
      'e2e', round(d['e2e']['ms_per_step'], 2), d['e2e']['value'], 'full', round(d['e2e_variants']['full_copy_back']['ms_per_step'], 2),
      'parity', d.get('parity', {}).get('max_rel'), d.get('parity', {}).get('rhs_max_rel'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
