cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/r02ac_bench.json 2> gpurun_out/r02ac_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02ac_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02ac_bench.json').read()); print(d['value'], d['ms_per_block'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac']); print(json.dumps(d['alt_ingest'])[:1500]); print(d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600
python bench.py --workload cfg5 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg5', d['value'], d['ms_per_block'], d['e2e']['value'], d['roofline']['frac'])"
python bench.py --workload cfg3 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg3', d['value'], d['ms_per_block'], d['e2e']['value'], d['roofline']['frac'])"
