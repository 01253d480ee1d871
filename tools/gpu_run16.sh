cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_hb3r|k_hb_tail|k_post_pre|k_post_fir|k_post_mid|k_resample|k_fastfir|k_post_seq' -s 30 -c 24 -o gpurun_out/r02o_k2_burst -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02o_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02o_k2_burst.ncu-rep gpurun_out/r02o_k2_burst_summary.csv
cut -d, -f1-6,9,12-14,16-20 gpurun_out/r02o_k2_burst_summary.csv | head -30
