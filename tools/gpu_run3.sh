set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_parity.py::test_stereo_output_through_the_bank_resampler tests/test_gpu_parity.py::test_stereo_output_paths "tests/test_gpu_fullsize.py::test_rates_whose_10ms_block_is_not_a_multiple_of_2_pow_stages" -m gpu -q -x -s > gpurun_out/r02c_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.txt
tail -30 gpurun_out/r02c_pytest.txt
