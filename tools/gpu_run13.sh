cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fp16_form" > gpurun_out/r02ab_fp16test.log 2>&1; grep -E "^E|Error|passed|failed" gpurun_out/r02ab_fp16test.log | head -12
python -m pytest tests -m gpu -x -q -k "agc or demodulator_chain or bank_mixed or golden or config4 or config5_slice or stereo or cdemodulator or wire_formats or retune or smeter or cfg3" 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02ab_bench.json').read()); print(d['value'], d['ms_per_block'], d['e2e']['value'], d['roofline']['launch_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02ab_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02ab_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r02ab_launches.csv > gpurun_out/r02ab_table.txt; head -12 gpurun_out/r02ab_table.txt
