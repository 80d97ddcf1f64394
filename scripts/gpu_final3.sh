#!/bin/bash
# final one-GPU record of round 2 (second half): full GPU test suite, default bench line, ncu launch list of the bench command,
# full captures of the right-hand-side kernels (CSV pages)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n1.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
for k in tp_rhs_cell_local_kernel tp_rhs_local_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/cap_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/cap_$k.log 2>&1; echo "capture $k rc=$?"
  ncu -i /tmp/cap_$k.ncu-rep --page raw --csv > gpurun_out/cap_$k.raw.csv 2>/dev/null
  ncu -i /tmp/cap_$k.ncu-rep --page source --csv > gpurun_out/cap_$k.src.csv 2>/dev/null
done
ls gpurun_out | head -40
