mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_6_2gpu.json 2> gpurun_out/r2/bench_6_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2/bench_6_2gpu_ref.json 2> gpurun_out/r2/bench_6_2gpu_ref.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload vision --steps 10 --warmup 3 > gpurun_out/r2/bench_6_2gpu_vision.json 2> gpurun_out/r2/bench_6_2gpu_vision.err
head -c 700 gpurun_out/r2/bench_6_2gpu.json; echo; tail -5 gpurun_out/r2/bench_6_2gpu.err; head -c 300 gpurun_out/r2/bench_6_2gpu_ref.json; echo; head -c 400 gpurun_out/r2/bench_6_2gpu_vision.json; tail -5 gpurun_out/r2/bench_6_2gpu_vision.err
