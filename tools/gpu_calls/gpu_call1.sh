mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2/gpu.txt
( python -m pip install mujoco dm_control 2>&1 | tail -3; python -m pip download mujoco -d /tmp/x 2>&1 | tail -2; ls /opt/wheelhouse | grep -i -E "mujoco|dm_control|dm-control|dm_env" ) > gpurun_out/r2/pip_probe.log 2>&1
python -m pytest tests -m gpu -q -rA > gpurun_out/r2/gpu_tests_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_1.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 4 > gpurun_out/r2/bench_0.json 2> gpurun_out/r2/bench_0.err
python bench.py --steps 20 --warmup 5 --no-cpu --envs 16384 > gpurun_out/r2/bench_0_16k.json 2> gpurun_out/r2/bench_0_16k.err
for tool in racecheck synccheck memcheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/gpu_sanitize.py 32 > gpurun_out/r2/sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_$tool.log
done
tail -3 gpurun_out/r2/gpu_tests_1.log; tail -2 gpurun_out/r2/sanitizer_*.log; cat gpurun_out/r2/pip_probe.log | tail -5
