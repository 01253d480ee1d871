# N-GPU bench (weak and strong scaling) through the library's NCCL broadcast.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh 8'   (outputs under gpurun_out/rNN_*)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
run() { # label, extra args
  label=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/rNN_bench_${N}gpu_$label.json 2> gpurun_out/rNN_${N}gpu_$label.err
  python -c "
import json; d=json.loads(open('gpurun_out/rNN_bench_${N}gpu_$label.json').read()); a=d.get('alt_ingest') or {}; print('$label', round(d['value']), d['ms_per_block'], round(d['e2e']['value']), 'alt', round(a.get('value',0)), round((a.get('e2e') or {}).get('value',0)))"
}
run weak
run strong --scaling strong
grep -c "NCCL" gpurun_out/rNN_${N}gpu_weak.err
