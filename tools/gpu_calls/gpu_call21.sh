mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --workload vision --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_21_2gpu_vision.json 2> gpurun_out/r2/bench_21_2gpu_vision.err
head -c 1200 gpurun_out/r2/bench_21_2gpu_vision.json; echo; tail -5 gpurun_out/r2/bench_21_2gpu_vision.err
