mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_9.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_9_ref.json 2> gpurun_out/r2/bench_9_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_9.json 2> gpurun_out/r2/bench_9.err
python tools/gpu_longrun.py 4096 300 0.3 > gpurun_out/r2/longrun4_4096_tcd03.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_memcheck_9.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_memcheck_9.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_racecheck_9.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_racecheck_9.log
tail -3 gpurun_out/r2/gpu_tests_9.log; head -c 300 gpurun_out/r2/bench_9.json; echo; head -c 300 gpurun_out/r2/bench_9_ref.json; echo; tail -1 gpurun_out/r2/longrun4_4096_tcd03.log | cut -c1-200; tail -2 gpurun_out/r2/sanitizer_memcheck_9.log gpurun_out/r2/sanitizer_racecheck_9.log
