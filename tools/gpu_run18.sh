cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests -m gpu -x -q -k "async or device_resident or two_banks or rewritten" 2>&1 | tail -2
run() { # label, env..., extra args
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python -c "
import json; d=json.loads(open('gpurun_out/tmp.json').read()); a=d.get('alt_ingest') or {}; print('$label', round(d['value']), d['ms_per_block'], round(d['e2e']['value']), 'alt', round(a.get('value',0)), round((a.get('e2e') or {}).get('value',0)))"
}
EXTRA=""
run slots4 X=1
run slots4_chunk8M CUTESDR_BCAST_CHUNK_KB=8192
