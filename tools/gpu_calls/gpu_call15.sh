mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_gather.so $V/libfb_fsched.so --rounds 2 > gpurun_out/r2/ab_fsched.log 2>&1
cut -c1-330 gpurun_out/r2/ab_fsched.log | tail -6
