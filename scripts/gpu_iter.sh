#!/bin/bash
# quick iteration: template tests, bench, and one full ncu capture of the kernels matching $1
mkdir -p gpurun_out
if [ -n "$SAN" ]; then timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "template_path" 2>&1 | tail -4; fi
timeout 600 python -m pytest tests -m gpu -x -q -k "${2:-template or fastpath or linear}" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ -n "$1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -c ${3:-1} -f -o gpurun_out/prof_iter \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_iter.log 2>&1
echo "ncu rc=$?"
fi
