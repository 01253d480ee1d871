set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_smi.txt 2>&1; nproc >> gpurun_out/r02a_smi.txt
python tests/tools/acq_probe.py > gpurun_out/r02a_acq.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
tail -3 gpurun_out/r02a_pytest.txt
