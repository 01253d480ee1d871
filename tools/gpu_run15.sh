cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for v in "X=1" "CUTESDR_HS_CTAS=2" "CUTESDR_HS_CTAS=1"; do
  echo "== $v"; env $v $B 2>>gpurun_out/r02p.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_block'], d['roofline']['launch_ms'], d['e2e']['value'])"
done
python -m pytest tests -m gpu -x -q -k "kernel2 or bit_ident or config4 or downconvert" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02p_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r02p_launches.csv 2>/dev/null | head -6
