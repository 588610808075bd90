mkdir -p gpurun_out/r2
V=flybody_b200/lib/variants
python tools/gpu_ab.py $V/libfb_base.so $V/libfb_mrow4.so $V/libfb_mrow2.so $V/libfb_rootonly.so --rounds 2 > gpurun_out/r2/ab_mrow.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2/smoke_20.log 2>&1
cut -c1-330 gpurun_out/r2/ab_mrow.log | tail -9; tail -2 gpurun_out/r2/smoke_20.log
