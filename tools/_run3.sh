set -x
timeout 300 python -m pytest tests/test_ba_gpu.py tests/test_corr_gpu.py -m gpu -q -k "cholesky or blocks or c3 or corr or golden or oracle" > gpurun_out/r02_chol_tests.log 2>&1; tail -5 gpurun_out/r02_chol_tests.log
for v in "" "VGG_CHOL_LOOKAHEAD=0"; do env $v timeout 120 python tools/microbench.py chol 2403; done > gpurun_out/r02_chol_bench.log 2>&1
for v in "" "VGG_CHOL=lib"; do env $v timeout 120 python tools/microbench.py ba; done >> gpurun_out/r02_chol_bench.log 2>&1
timeout 120 python tools/microbench.py blocks 4096 >> gpurun_out/r02_chol_bench.log 2>&1
timeout 120 python tools/microbench.py blocks 131072 >> gpurun_out/r02_chol_bench.log 2>&1
cat gpurun_out/r02_chol_bench.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 1500 gpurun_out/r02_bench_b.json; tail -5 gpurun_out/r02_bench_b.err
