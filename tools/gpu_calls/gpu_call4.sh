mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
( cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us /sys/fs/cgroup/cpu/cpu.cfs_period_us; nproc; lscpu | head -20 ) > gpurun_out/r2/cpuinfo.txt 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_4.log
python tools/gpu_ab.py $V/libfb_D.so $V/libfb_E.so $V/libfb_F.so --envs 4096,16384 > gpurun_out/r2/ab_DEF.log 2>&1
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/r2/bench_3.json 2> gpurun_out/r2/bench_3.err
python bench.py --workload vision --steps 20 --warmup 5 > gpurun_out/r2/bench_3_vision.json 2> gpurun_out/r2/bench_3_vision.err
python tools/gpu_longrun.py 4096 500 0.3 > gpurun_out/r2/longrun2_4096_tcd03.log 2>&1
timeout 700 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_racecheck_4.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_racecheck_4.log
tail -3 gpurun_out/r2/gpu_tests_4.log; grep SUMMARY gpurun_out/r2/ab_DEF.log; tail -2 gpurun_out/r2/sanitizer_racecheck_4.log; head -c 400 gpurun_out/r2/bench_3_vision.json; tail -3 gpurun_out/r2/bench_3_vision.err; cat gpurun_out/r2/cpuinfo.txt | head -8
