cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for v in "X=1" "CUTESDR_TC_SPARE=8" "CUTESDR_TC_SPARE=12" "CUTESDR_TC_SPARE=20" "CUTESDR_TC_SPARE=28"; do
  echo "== $v"; env $v $B 2>>gpurun_out/r02m.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_block'], d['roofline']['launch_ms'], d['e2e']['value'])"
done
