mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_chol0.so $V/libfb_chol16.so $V/libfb_chol24.so $V/libfb_chol32.so --rounds 2 > gpurun_out/r2/ab_chol.log 2>&1
cut -c1-420 gpurun_out/r2/ab_chol.log | tail -12
