#!/bin/bash
# bench under several EXTFEM_OPTIONS settings: scripts/gpu_opts.sh "opt1=v,opt2=v" "opt..." ...   (phase_ms of each)
mkdir -p gpurun_out
for o in "$@"; do
  echo "== $o"
  EXTFEM_OPTIONS="$o" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: round(v, 4) for k, v in d['phase_ms'].items()}, 'step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'chk', d.get('checksum_sum_nzval'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done
