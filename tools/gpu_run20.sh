cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fp16_form or wire_formats" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.max,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum --clock-control none -k regex:'k_post_seq' -s 4 -c 4 --csv --log-file gpurun_out/r02u_seq.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02u_seq.csv',errors='ignore')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
for r in rows[hdr+1:]:
    if len(r)>vi: print(r[0], r[ki][:24], r[mi], r[vi])
PY
