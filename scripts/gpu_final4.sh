#!/bin/bash
# final build: ncu launch list of the bench command, configs 3 and 4
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
for c in 3 4; do
  timeout 600 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"; tail -c 1200 gpurun_out/config$c.json; echo
done
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1.json'))
print({k: round(v, 4) for k, v in d['phase_ms'].items()}, 'step', round(d['ms_per_step'], 4), 'value', d['value'], 'frac', round(d['roofline']['frac'], 4),
      'e2e', round(d['e2e']['ms_per_step'], 2), d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], 'full', round(d['e2e_variants']['full_copy_back']['ms_per_step'], 2),
      'parity', d.get('parity', {}).get('max_rel'), d.get('parity', {}).get('rhs_max_rel'), 'e2e_vs_resident', d.get('e2e_vs_resident_max_rel'),
      'resident', d['e2e_variants']['resident_solve']['ms_coords_in_to_system_ready'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
