set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=10 -s > gpurun_out/r02b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.txt
grep -E "fullsize\]|passed|failed" gpurun_out/r02b_pytest.txt | tail -20
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02b_bench_plain.json 2> gpurun_out/r02b_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file gpurun_out/r02b_launches.csv $B > gpurun_out/r02b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_post|k_fastfir|k_hb11|k_halfband|k_resample|k_fir' -s 300 -c 24 -o gpurun_out/r02b_burst $B > gpurun_out/r02b_ncu2.log 2>&1
ls -la gpurun_out | tail -8
