mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_33.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2/smoke_33.log 2>&1
tail -3 gpurun_out/r2/gpu_tests_33.log; tail -1 gpurun_out/r2/smoke_33.log
