cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_mix_tc' -s 6 -c 2 -o gpurun_out/r02ae_k1t_tf32 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alt-ingest > gpurun_out/r02ae_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_mix_tc' -s 6 -c 2 -o gpurun_out/r02ae_k1t_fp16 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alt-ingest --ingest cs16 > gpurun_out/r02ae_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:'k_nb|k_scan|k_dispfft|k_screen|k_unpack|k_plot' -c 24 -o gpurun_out/r02ae_cfg5_aux -f python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02ae_ncu3.log 2>&1
for f in k1t_tf32 k1t_fp16 cfg5_aux; do python tools/ncu_summary.py gpurun_out/r02ae_$f.ncu-rep gpurun_out/r02ae_${f}_summary.csv; done
cut -d, -f1-6,10-13,16-20 gpurun_out/r02ae_k1t_tf32_summary.csv gpurun_out/r02ae_k1t_fp16_summary.csv
cut -d, -f1-6,12,16-18 gpurun_out/r02ae_cfg5_aux_summary.csv | head -30
