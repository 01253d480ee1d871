cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_post_seq' -s 2 -c 2 -o gpurun_out/r02j_seq -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02j_ncu_seq.log 2>&1
ls -la gpurun_out/r02j_seq.ncu-rep
