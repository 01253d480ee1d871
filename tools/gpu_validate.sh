# One gpurun call: the whole -m gpu suite, the default bench line, and the ncu launch list of one bench step.
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'   (outputs under gpurun_out/rNN_*)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/rNN_gputests.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/rNN_gputests.log
python bench.py > gpurun_out/rNN_bench.json 2> gpurun_out/rNN_bench.err; echo "bench exit $?"
cat gpurun_out/rNN_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/rNN_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/rNN_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/rNN_launches.csv',errors='ignore')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)>vi:
        d[r[ki][:60]][0]+=1; d[r[ki][:60]][1]+=float(r[vi].replace(',',''))
tot=sum(v[1] for v in d.values())
for k,v in sorted(d.items(),key=lambda kv:-kv[1][1])[:25]: print("%-62s %4d %10.1f us %5.1f%%"%(k,v[0],v[1]/1e3,100*v[1]/tot))
PY
