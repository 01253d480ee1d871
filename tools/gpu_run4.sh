set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=10 -x > gpurun_out/r02d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.txt
tail -5 gpurun_out/r02d_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?" >> gpurun_out/r02d_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --ingest cs16 --no-cpu-baseline > gpurun_out/r02d_bench_cs16.json 2> gpurun_out/r02d_bench_cs16.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/r02d_bench_cfg5.json 2> gpurun_out/r02d_bench_cfg5.err
cat gpurun_out/r02d_bench*.json | cut -c1-600
