mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_fsched.so $V/libfb_froot.so --rounds 2 > gpurun_out/r2/ab_froot.log 2>&1
cut -c1-330 gpurun_out/r2/ab_froot.log | tail -6
