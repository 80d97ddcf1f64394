#!/bin/bash
# full ncu captures (one launch each) of the nonlinear-path kernels on config ${1:-4}; exported on the box as the raw and source
# CSV pages (the reports themselves exceed the transfer limit): gpurun_out/nl_<kernel>_c<config>.{raw,src}.csv
mkdir -p gpurun_out
c=${1:-4}
shift
for k in "${@:-local_nonlinear_kernel4}"; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -f -o /tmp/nl_${k}_c$c \
     python bench_configs.py $c > gpurun_out/ncufull_c$c.log 2>&1
  echo "$k rc=$?"
  ncu -i /tmp/nl_${k}_c$c.ncu-rep --page raw --csv > gpurun_out/nl_${k}_c$c.raw.csv 2>/dev/null
  ncu -i /tmp/nl_${k}_c$c.ncu-rep --page source --csv > gpurun_out/nl_${k}_c$c.src.csv 2>/dev/null
done
ls -la gpurun_out/
