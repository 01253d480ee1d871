cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fastfir or demodulator_chain or bank_mixed or golden or config4 or config3" 2>&1 | tail -5
python bench.py --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench.json').read()); print(d['value'], d['ms_per_block'], d['e2e']['value'], d['roofline']['launch_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r02n_launches.csv | head -12
