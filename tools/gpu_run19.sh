cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernel1t.py tests/test_gpu_ingest.py -m gpu -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "int16" 2>&1 | tail -5
for ing in cf32 cs16; do
python bench.py --ingest $ing --no-cpu-baseline 2>gpurun_out/r02r_$ing.err | tee gpurun_out/r02r_bench_$ing.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$ing', round(d['value']), d['ms_per_block'], round(d['e2e']['value']), d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline'].get('kernel'))"
done
