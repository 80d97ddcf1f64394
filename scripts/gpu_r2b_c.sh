#!/bin/bash
# round 2, second half, run C: full ncu captures of the two kernels of the cell-local right-hand side (CSV pages)
mkdir -p gpurun_out
for k in tp_rhs_cell_local_kernel tp_rhs_local_kernel; do
  EXTFEM_OPTIONS="$1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/cap_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/cap_$k.log 2>&1; echo "capture $k rc=$?"
  ncu -i /tmp/cap_$k.ncu-rep --page raw --csv > gpurun_out/cap_$k.raw.csv 2>/dev/null
  ncu -i /tmp/cap_$k.ncu-rep --page source --csv > gpurun_out/cap_$k.src.csv 2>/dev/null
  ncu -i /tmp/cap_$k.ncu-rep --page details > gpurun_out/cap_$k.details.txt 2>/dev/null
done
ls -la gpurun_out | head -30
