cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # label, env..., extra args
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python -c "
import json; d=json.loads(open('gpurun_out/tmp.json').read()); print('$label', round(d['value']), d['ms_per_block'], round(d['e2e']['value']))"
}
EXTRA=""
run chunk4M X=1
run chunk8M CUTESDR_BCAST_CHUNK_KB=8192
run chunk2M CUTESDR_BCAST_CHUNK_KB=2048
EXTRA="--ingest cs16"
run cs16_chunk4M X=1
run cs16_chunk2M CUTESDR_BCAST_CHUNK_KB=2048
EXTRA="--scaling strong"
run strong_cf32 X=1
EXTRA="--scaling strong --ingest cs16"
run strong_cs16 X=1
