#!/bin/bash
# GPU pass: parity tests, bench line, ncu launch list.  $1 = "ncu" adds the launch list, $2 = "san" runs the
# template tests under compute-sanitizer first.
mkdir -p gpurun_out
if [ "$2" == "san" ]; then
timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "template_path" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -25 gpurun_out/sanitizer.log
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"
fi
