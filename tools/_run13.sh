set -x
timeout 120 python tools/syrk_i8_check.py rate > gpurun_out/r02_bench13.log 2>&1
for ts in 1 0; do
  VGG_SYRK_TS=$ts timeout 200 python tools/syrk_i8_check.py 640 1280 7 >> gpurun_out/r02_bench13.log 2>&1
  VGG_SYRK_TS=$ts timeout 300 python tools/syrk_i8_check.py 2432 12288 7 time >> gpurun_out/r02_bench13.log 2>&1
  VGG_SYRK_TS=$ts timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench13.log 2>&1
done
timeout 120 python tools/microbench.py chol128 >> gpurun_out/r02_bench13.log 2>&1
timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench13.log 2>&1
cat gpurun_out/r02_bench13.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t13.log 2>&1; tail -5 gpurun_out/r02_t13.log
VGG_SYRK_TS=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t13b.log 2>&1; tail -3 gpurun_out/r02_t13b.log
