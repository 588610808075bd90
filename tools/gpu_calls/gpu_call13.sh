mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_13.log
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 363 --launch-count 8 -f -o gpurun_out/r2/prof_v10 python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v10.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/launches_bench_v10.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --preroll 0 > gpurun_out/r2/bench_under_ncu_v10.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_13.json 2> gpurun_out/r2/bench_13.err
tail -3 gpurun_out/r2/gpu_tests_13.log; tail -3 gpurun_out/r2/ncu_v10.log; head -c 400 gpurun_out/r2/bench_13.json; ls -la gpurun_out/r2/prof_v10.ncu-rep
