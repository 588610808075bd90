mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_34.log
tail -3 gpurun_out/r2/gpu_tests_34.log
