mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench_22_8gpu.json 2> gpurun_out/r2/bench_22_8gpu.err
head -c 900 gpurun_out/r2/bench_22_8gpu.json; echo; tail -3 gpurun_out/r2/bench_22_8gpu.err
