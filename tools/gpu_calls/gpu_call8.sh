mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_8.log
python tools/gpu_ab.py $V/libfb_G.so $V/libfb_H.so $V/libfb_H.so:FB_NO_HEAVY_KERNEL=1 --envs 4096 > gpurun_out/r2/ab_GH.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2/bench_7.json 2> gpurun_out/r2/bench_7.err
FB_NO_HEAVY_KERNEL=1 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2/bench_7_noheavy.json 2> gpurun_out/r2/bench_7_noheavy.err
python tools/gpu_longrun.py 4096 300 0.3 > gpurun_out/r2/longrun3_4096_tcd03.log 2>&1
FB_BENCH_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_8_2gpu.json 2> gpurun_out/r2/bench_8_2gpu.err
tail -3 gpurun_out/r2/gpu_tests_8.log; grep SUMMARY gpurun_out/r2/ab_GH.log; head -c 250 gpurun_out/r2/bench_7.json; echo; head -c 250 gpurun_out/r2/bench_7_noheavy.json; echo; tail -1 gpurun_out/r2/longrun3_4096_tcd03.log | cut -c1-200; head -c 250 gpurun_out/r2/bench_8_2gpu.json; grep "rank" gpurun_out/r2/bench_8_2gpu.err | tail -4
