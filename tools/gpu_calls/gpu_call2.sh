mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python -m pytest tests -m gpu -q -x > gpurun_out/r2/gpu_tests_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_2.log
python tools/gpu_ab.py $V/libfb_A.so $V/libfb_B.so $V/libfb_C.so --envs 4096,16384 > gpurun_out/r2/ab_ABC.log 2>&1
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2/bench_1.json 2> gpurun_out/r2/bench_1.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --preroll 0 > gpurun_out/r2/bench_1_standing.json 2> gpurun_out/r2/bench_1_standing.err
python tools/gpu_longrun.py 4096 1000 > gpurun_out/r2/longrun_4096.log 2>&1
python tools/gpu_longrun.py 4096 500 0.3 > gpurun_out/r2/longrun_4096_tcd03.log 2>&1
timeout 300 python tools/gpu_vision.py 1024 30 > gpurun_out/r2/vision_1.log 2>&1
timeout 420 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_sanitize.py 32 > gpurun_out/r2/sanitizer_racecheck_2.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_racecheck_2.log
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 357 --launch-count 8 -f -o gpurun_out/r2/prof_v8 python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v8.log 2>&1
tail -3 gpurun_out/r2/gpu_tests_2.log; grep SUMMARY gpurun_out/r2/ab_ABC.log; tail -2 gpurun_out/r2/longrun_4096.log | cut -c1-300; tail -3 gpurun_out/r2/vision_1.log; tail -2 gpurun_out/r2/sanitizer_racecheck_2.log; ls -la gpurun_out/r2/
