#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "template" 2>&1 | tail -2
for o in "template_pool_bytes=37376" "template_pool_bytes=31744" "template_pool_bytes=30720" "template_pool_bytes=36000"; do
  echo "== $o"
  EXTFEM_OPTIONS=$o python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'], d['plan']['template_ctas'])"
done
