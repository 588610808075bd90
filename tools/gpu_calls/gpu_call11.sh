mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -k "vision" > gpurun_out/r2/gpu_tests_11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_11.log
python bench.py --workload vision --steps 30 --warmup 5 > gpurun_out/r2/bench_11_vision.json 2> gpurun_out/r2/bench_11_vision.err
tail -3 gpurun_out/r2/gpu_tests_11.log; head -c 1500 gpurun_out/r2/bench_11_vision.json; echo; tail -5 gpurun_out/r2/bench_11_vision.err
