#!/bin/bash
# per-kernel launch durations of configs 3 and 4 (ncu, time metric only) -> gpurun_out/nl_launches_c{3,4}.csv
mkdir -p gpurun_out
for c in 4 3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nonlinear|nl_point|gather_columns|gather' -c 40 --csv \
    --log-file gpurun_out/nl_launches_c$c.csv python bench_configs.py $c > gpurun_out/ncu_c$c.log 2>&1
  echo "config $c rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/nl_launches_c$c.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size'); ib=h.index('Block Size')
for r in rows[1:13]: print(r[ik][:70], r[ig], r[ib], r[iv])
PY
done
