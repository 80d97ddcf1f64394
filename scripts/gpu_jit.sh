#!/bin/bash
mkdir -p gpurun_out
EXTFEM_JIT_VERBOSE=1 timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "specialised" > gpurun_out/jit_test.log 2>&1; echo "jit test rc=$?"; tail -15 gpurun_out/jit_test.log
EXTFEM_OPTIONS=template_jit=1 EXTFEM_JIT_VERBOSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['phase_ms'], d['roofline']['frac'], d['setup_s'], d['checksum_sum_nzval'])"
EXTFEM_OPTIONS=template_jit=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('static', d['ms_per_step'], d['phase_ms'], d['checksum_sum_nzval'])"
