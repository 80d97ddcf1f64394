#!/bin/bash
# configs 3 and 4 under several EXTFEM_OPTIONS settings (ms per assembly and phases): scripts/gpu_nl_opts.sh "k=v,k=v" ...
mkdir -p gpurun_out
for o in "$@"; do
  for c in 3 4; do
    EXTFEM_OPTIONS="$o" EXTFEM_NO_PARITY=1 timeout 300 python bench_configs.py $c 2> gpurun_out/cfg.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$o', 'config', d['config'], round(d['ms'], 3), {k: round(v, 3) for k, v in d['phase_ms'].items()}, d['checks']['ok'])
"
  done
done
