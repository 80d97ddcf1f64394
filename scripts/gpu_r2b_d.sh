#!/bin/bash
# round 2, second half, run D: right-hand-side variants after the first profile (constant-bank coefficients, compile-time local
# dofs, occupancy variants of the gather)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2b.py -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "r2b pytest rc=$?"; tail -15 gpurun_out/pytest_r2b.log
first=1
for o in "rhs_groups=1" "rhs_gather_ctas=6" "rhs_gather_ctas=8" "rhs_fast_trig=0"; do
  echo "== $o"
  EXTFEM_OPTIONS="$o" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $([ $first = 1 ] || echo --no-parity) > gpurun_out/bench_opt.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  first=0
  cp gpurun_out/bench_opt.json "gpurun_out/bench_d_$(echo $o | tr ',=' '__').json"
  python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_opt.json'))
print({k: round(v, 4) for k, v in d['phase_ms'].items()}, 'step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4),
      'e2e', round(d['e2e']['ms_per_step'], 2), 'parity', d.get('parity', {}).get('max_rel'), d.get('parity', {}).get('rhs_max_rel'),
      'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
