cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_hb3r|k_hb_tail' -s 20 -c 4 -o gpurun_out/r02g_k2 -f $B > gpurun_out/r02g_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02g_k2.ncu-rep gpurun_out/r02g_k2_summary.csv
cat gpurun_out/r02g_k2_summary.csv
for v in "CUTESDR_HS_CTAS=1" "CUTESDR_HS_CTAS=2" ; do
  echo "== $v"; env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/r02e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'])"
done
