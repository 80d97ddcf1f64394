#!/bin/bash
# one full ncu capture of the tensor-core nonlinear kernel on config ${1:-4}
mkdir -p gpurun_out
c=${1:-4}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${2:-local_nonlinear_kernel4}" -s 1 -c 1 -f -o gpurun_out/nl4_c$c \
   python bench_configs.py $c > gpurun_out/ncufull_c$c.log 2>&1
echo "rc=$?"; ls -la gpurun_out/nl4_c$c.ncu-rep
