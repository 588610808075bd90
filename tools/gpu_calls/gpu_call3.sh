mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_3.log
python tools/gpu_ab.py $V/libfb_B.so $V/libfb_D.so --envs 4096,16384 > gpurun_out/r2/ab_BD.log 2>&1
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2/bench_2.json 2> gpurun_out/r2/bench_2.err
timeout 300 python tools/gpu_vision.py 1024 30 > gpurun_out/r2/vision_2.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_sanitize.py 32 > gpurun_out/r2/sanitizer_racecheck_3.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_racecheck_3.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 30 python tools/gpu_sanitize.py 32 > gpurun_out/r2/sanitizer_memcheck_3.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_memcheck_3.log
tail -3 gpurun_out/r2/gpu_tests_3.log; grep SUMMARY gpurun_out/r2/ab_BD.log; tail -4 gpurun_out/r2/vision_2.log; tail -2 gpurun_out/r2/sanitizer_racecheck_3.log gpurun_out/r2/sanitizer_memcheck_3.log; head -c 600 gpurun_out/r2/bench_2.json; tail -3 gpurun_out/r2/bench_2.err
