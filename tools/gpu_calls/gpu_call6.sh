mkdir -p gpurun_out/r2
timeout 300 python tools/gpu_vision_profile.py 1024 > gpurun_out/r2/vision_profile.log 2>&1
python bench.py --workload vision --steps 20 --warmup 5 > gpurun_out/r2/bench_5_vision.json 2> gpurun_out/r2/bench_5_vision.err
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 363 --launch-count 8 -f -o gpurun_out/r2/prof_v9 python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v9.log 2>&1
FB_VARIANT=flight timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 147 --launch-count 8 -f -o gpurun_out/r2/prof_v9_flight python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v9_flight.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --preroll 0 > gpurun_out/r2/bench_under_ncu.log 2>&1
head -50 gpurun_out/r2/vision_profile.log; head -c 300 gpurun_out/r2/bench_5_vision.json; tail -3 gpurun_out/r2/ncu_v9.log gpurun_out/r2/ncu_v9_flight.log
