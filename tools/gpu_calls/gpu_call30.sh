mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/gpu_tests_30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gpu_tests_30.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2/smoke_30.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 363 --launch-count 8 -f -o gpurun_out/r2/prof_v13 python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v13.log 2>&1
FB_VARIANT=flight timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 147 --launch-count 8 -f -o gpurun_out/r2/prof_v13_flight python tools/gpu_ncu_target.py > gpurun_out/r2/ncu_v13_flight.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/launches_bench_v13.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --preroll 0 > gpurun_out/r2/bench_under_ncu_v13.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_30_ref.json 2> gpurun_out/r2/bench_30_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_30.json 2> gpurun_out/r2/bench_30.err
python tools/gpu_longrun.py 4096 300 0.3 > gpurun_out/r2/longrun_30.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_memcheck_30.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_memcheck_30.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_racecheck_30.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_racecheck_30.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 30 python tools/gpu_sanitize.py 16 > gpurun_out/r2/sanitizer_synccheck_30.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sanitizer_synccheck_30.log
tail -3 gpurun_out/r2/gpu_tests_30.log; tail -1 gpurun_out/r2/smoke_30.log; head -c 300 gpurun_out/r2/bench_30.json; echo; head -c 300 gpurun_out/r2/bench_30_ref.json; echo; tail -1 gpurun_out/r2/longrun_30.log | cut -c1-200
for f in memcheck racecheck synccheck; do tail -2 gpurun_out/r2/sanitizer_${f}_30.log; done
