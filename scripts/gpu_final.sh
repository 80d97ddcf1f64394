#!/bin/bash
# Final 1-GPU pass of a round: full GPU test suite, bench (ours + reference arm), ncu launch list and full captures of the
# hot kernels, full-size runs of configs 3 and 4.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fp_geo_kernel|tp_gather_kernel|tp_rhs_cell_kernel|tp_rhs_kernel" -s 4 -c 4 -f -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_final.log 2>&1; echo "ncu rc=$?"
timeout 500 python bench_configs.py 3 > gpurun_out/config3.json 2> gpurun_out/config3.err; echo rc3=$?; cat gpurun_out/config3.json
timeout 500 python bench_configs.py 4 > gpurun_out/config4.json 2> gpurun_out/config4.err; echo rc4=$?; cat gpurun_out/config4.json
