cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "rewritten or two_banks" 2>&1 | tail -3
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --scaling $sc --no-cpu-baseline > gpurun_out/r02q_bench_2gpu_$sc.json 2> gpurun_out/r02q_2gpu_$sc.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_2gpu_$sc.json').read()); print('$sc', d['value'], d['ms_per_block'], d['e2e'])"
done
python bench.py --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu', d['value'], d['ms_per_block'], d['e2e']['value'])"
grep -i "nccl" gpurun_out/r02q_2gpu_weak.err | head -5
