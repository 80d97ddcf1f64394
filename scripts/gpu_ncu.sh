#!/bin/bash
# full ncu capture of the kernels named by regex $1: $3 launches after skipping $4
PAT=${1:-fp_gather_kernel}
OUT=${2:-prof_gather}
CNT=${3:-1}
SKIP=${4:-0}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c $CNT -f -o gpurun_out/$OUT \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${OUT}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${OUT}.log
