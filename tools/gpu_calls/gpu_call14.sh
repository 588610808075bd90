mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_chol24.so $V/libfb_gather.so $V/libfb_gather2.so --rounds 2 > gpurun_out/r2/ab_gather.log 2>&1
cut -c1-330 gpurun_out/r2/ab_gather.log | tail -8
