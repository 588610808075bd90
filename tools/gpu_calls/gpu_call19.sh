mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_19_2gpu.json 2> gpurun_out/r2/bench_19_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench_19_2gpu_ref.json 2> gpurun_out/r2/bench_19_2gpu_ref.err
head -c 600 gpurun_out/r2/bench_19_2gpu.json; echo; head -c 300 gpurun_out/r2/bench_19_2gpu_ref.json; echo; tail -5 gpurun_out/r2/bench_19_2gpu.err
