set -x
timeout 400 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "c3 or cholesky or global_ba" > gpurun_out/r02_t8.log 2>&1; tail -6 gpurun_out/r02_t8.log
timeout 120 python tools/microbench.py chol 2403 > gpurun_out/r02_bench8.log 2>&1
timeout 120 python tools/microbench.py blocks 4096 >> gpurun_out/r02_bench8.log 2>&1
timeout 120 python tools/microbench.py blocks 131072 >> gpurun_out/r02_bench8.log 2>&1
timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench8.log 2>&1
cat gpurun_out/r02_bench8.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol2.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_lm.csv python tools/microbench.py ba > /dev/null 2>&1
