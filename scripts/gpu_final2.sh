#!/bin/bash
# nonlinear-path record of a build: configs 3/4/5 JSON, per-kernel launch times, full ncu captures (CSV pages) of the three
# nonlinear-path kernels on config 4 and of the contraction kernel on config 3
mkdir -p gpurun_out
bash scripts/gpu_configs.sh > gpurun_out/configs.log 2>&1; grep -E "rc=" gpurun_out/configs.log
for c in 4 3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nonlinear|nl_point|gather' -c 8 --csv \
    --log-file gpurun_out/nl_launches_c$c.csv python bench_configs.py $c > gpurun_out/ncu_c$c.log 2>&1
done
EXTFEM_NO_PARITY=1 bash scripts/gpu_nl_ncufull.sh 4 local_nonlinear_kernel4 nl_point gather_columns_warp > gpurun_out/ncufull4.log 2>&1
EXTFEM_NO_PARITY=1 bash scripts/gpu_nl_ncufull.sh 3 local_nonlinear_kernel4 > gpurun_out/ncufull3.log 2>&1
ls gpurun_out | head -50
