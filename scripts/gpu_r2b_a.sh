#!/bin/bash
# round 2, second half: parity tests of the cell-local right-hand side and the flat write-out, full GPU suite, A/B of the two
# options on the headline bench inside one build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2b.py -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "r2b pytest rc=$?"; tail -15 gpurun_out/pytest_r2b.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for o in "rhs_local=0" "rhs_local=1" "template_flat_writeout=20" "template_flat_writeout=30" "template_flat_writeout=1000" "rhs_local=0,template_flat_writeout=0"; do
  echo "== $o"
  EXTFEM_OPTIONS="$o" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_opt.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cp gpurun_out/bench_opt.json "gpurun_out/bench_$(echo $o | tr ',=' '__').json"
  python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_opt.json'))
print({k: round(v, 4) for k, v in d['phase_ms'].items()}, 'step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4),
      'e2e', round(d['e2e']['ms_per_step'], 2), 'parity', d.get('parity', {}).get('max_rel'), d.get('parity', {}).get('rhs_max_rel'),
      'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
