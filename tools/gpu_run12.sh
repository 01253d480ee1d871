cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/bin/seq_probe | tee gpurun_out/r02i_seq_probe.txt
for w in "--ingest cs16" "--workload cfg5" "--workload cfg3" "--workload cfg1" "--workload cfg2"; do
  n=$(echo $w | tr -d ' -' )
  python bench.py $w --no-cpu-baseline > gpurun_out/r02i_bench_$n.json 2>gpurun_out/r02i_err_$n.txt; echo "$w exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench_$n.json').read())
print(d['value'], d['ms_per_step'], d.get('ms_per_block'), d['e2e']['value'], d['roofline'].get('bound'), d['roofline'].get('frac'), d['roofline'].get('launch_ms'))
PY
done
