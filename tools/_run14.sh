set -x
rm -f gpurun_out/r02_bench14.log
for ts in 1 0; do
  VGG_SYRK_TS=$ts timeout 200 python tools/syrk_i8_check.py 640 1280 7 >> gpurun_out/r02_bench14.log 2>&1
  VGG_SYRK_TS=$ts timeout 300 python tools/syrk_i8_check.py 2432 12288 7 time >> gpurun_out/r02_bench14.log 2>&1
  VGG_SYRK_TS=$ts timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench14.log 2>&1
done
timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench14.log 2>&1
grep -v "^  File\|^    " gpurun_out/r02_bench14.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t14.log 2>&1; tail -4 gpurun_out/r02_t14.log
VGG_SYRK_TS=1 timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_syrk_i8_gpu.py -m gpu -q -x > gpurun_out/r02_t14b.log 2>&1; tail -4 gpurun_out/r02_t14b.log
timeout 300 python tools/video_c5.py --frames 160 --new 128 -v > gpurun_out/r02_c5_small.log 2>&1; tail -12 gpurun_out/r02_c5_small.log | cut -c1-600
