#!/bin/bash
# BASELINE configs 3, 4, 5 at full size on one GPU -> gpurun_out/config{3,4,5}.json
mkdir -p gpurun_out
for c in 3 4 5; do
  timeout 900 python bench_configs.py $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c rc=$?"
  tail -c 1800 gpurun_out/config$c.json; echo; tail -3 gpurun_out/config$c.err
done
