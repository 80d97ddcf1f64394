#!/bin/bash
# one-GPU record of a build: full GPU test suite, default bench line, reference arm, ncu launch list of the bench command and a
# full capture of the dominant kernels (exported as CSV pages; the reports exceed the transfer limit)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
for k in tw_gather_kernel tp_gather_kernel geo; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/cap_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/cap_$k.log 2>&1; echo "capture $k rc=$?"
  ncu -i /tmp/cap_$k.ncu-rep --page raw --csv > gpurun_out/cap_$k.raw.csv 2>/dev/null
  ncu -i /tmp/cap_$k.ncu-rep --page source --csv > gpurun_out/cap_$k.src.csv 2>/dev/null
done
ls -la gpurun_out | head -40
