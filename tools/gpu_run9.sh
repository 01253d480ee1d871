cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for v in "X=1" "CUTESDR_TC_SPARE=8" "CUTESDR_TC_SPARE=12" "CUTESDR_TC_SPARE=20" "CUTESDR_HS_CTAS=2" "CUTESDR_NO_HBTAIL=1"; do
  echo "== $v"; env $v $B 2>>gpurun_out/r02e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'])"
done
tail -5 gpurun_out/r02e.err
