cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fp16_form or wire_formats" > gpurun_out/r02x_fp16test.log 2>&1; grep -E "^E|Error|passed|failed" gpurun_out/r02x_fp16test.log | head -12
ncu --set full --clock-control none --import-source on -k regex:'k_post_seq2' -s 2 -c 1 -o gpurun_out/r02x_seq2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02x_ncu_seq.log 2>&1
ls -la gpurun_out/r02x_seq2.ncu-rep
