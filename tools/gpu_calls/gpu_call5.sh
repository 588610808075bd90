mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_F.so $V/libfb_G.so $V/libfb_G.so:FB_OVERLAP_VEL=1 --envs 4096,16384 > gpurun_out/r2/ab_FG.log 2>&1
FB_OVERLAP_VEL=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2/gpu_tests_5_overlap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_5_overlap.log
FB_OVERLAP_VEL=1 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2/bench_4_overlap.json 2> gpurun_out/r2/bench_4_overlap.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2/bench_4.json 2> gpurun_out/r2/bench_4.err
grep SUMMARY gpurun_out/r2/ab_FG.log; tail -3 gpurun_out/r2/gpu_tests_5_overlap.log; head -c 300 gpurun_out/r2/bench_4_overlap.json; echo; head -c 300 gpurun_out/r2/bench_4.json
