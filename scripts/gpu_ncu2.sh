#!/bin/bash
# full ncu capture: kernels matching $1 into gpurun_out/$2 (EXTFEM_OPTIONS from $3)
mkdir -p gpurun_out
EXTFEM_OPTIONS="$3" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -c ${4:-1} -f -o gpurun_out/$2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/$2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/$2.log
