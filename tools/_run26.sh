set -x
timeout 120 python tools/microbench.py trsv 2402 > gpurun_out/r02_trsv_probe2.log 2>&1; head -8 gpurun_out/r02_trsv_probe2.log; tail -3 gpurun_out/r02_trsv_probe2.log
timeout 120 python tools/microbench.py ba > gpurun_out/r02_bench26.log 2>&1; cat gpurun_out/r02_bench26.log
timeout 400 python -m pytest tests/test_ba_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r02_t26.log 2>&1; tail -3 gpurun_out/r02_t26.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:trsv_upper -c 20 --csv --log-file gpurun_out/r02_launches_trsv2.csv python tools/microbench.py ba > /dev/null 2>&1; tail -2 gpurun_out/r02_launches_trsv2.csv | cut -c200-330
