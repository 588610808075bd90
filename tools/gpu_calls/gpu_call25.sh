mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_p6.so $V/libfb_p7.so --rounds 2 > gpurun_out/r2/ab_p7.log 2>&1
grep SUMMARY gpurun_out/r2/ab_p7.log
grep -o '"lib": "[a-z0-9_.]*", "envs": 4096, "round": 1, "ms_per_step": [0-9.]*, "env_steps_per_s": [0-9]*, "stages_ms": {[^}]*}' gpurun_out/r2/ab_p7.log
