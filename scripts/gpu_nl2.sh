#!/bin/bash
# nonlinear path: full GPU tests of the generic path, then configs 3 and 4 with per-kernel times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "${NLK:-nonlinear or neohooke or rcd or stvenant or example or generic or bilinear or block or stokes or mixed}" > gpurun_out/pytest_nl.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_nl.log
for c in 4 3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nonlinear|nl_point|gather' -c 8 --csv \
    --log-file gpurun_out/nl_launches_c$c.csv python bench_configs.py $c > gpurun_out/ncu_c$c.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/nl_launches_c$c.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size'); ib=h.index('Block Size')
for r in rows[5:9]: print(r[ik][:60], r[ig], r[ib], r[iv])
PY
  timeout 600 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/config$c.json')); print({k:d[k] for k in ('config','ms','phase_ms','checks')})"; tail -3 gpurun_out/config$c.err
done
