mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "dist rc=$?"; tail -2 gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; wc -l gpurun_out/bench_n2.json; head -c 200 gpurun_out/bench_n2.json; echo
EXTFEM_OPTIONS=fastpath_templates=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_records.json 2> gpurun_out/bench_records.err; echo "records rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_records.json')); print(d['ms_per_step'], d['phase_ms'], d['plan'], d['roofline']['frac'])"
