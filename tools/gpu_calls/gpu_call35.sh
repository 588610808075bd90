mkdir -p gpurun_out/r2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench_35_2gpu.json 2> gpurun_out/r2/bench_35_2gpu.err
head -c 400 gpurun_out/r2/bench_35_2gpu.json; echo
