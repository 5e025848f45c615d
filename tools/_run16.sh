set -x
rm -f gpurun_out/r02_bench16.log
for sp in 1 0; do
  VGG_CHOL_SPLIT=$sp timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench16.log 2>&1
  VGG_CHOL_SPLIT=$sp timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench16.log 2>&1
done
cat gpurun_out/r02_bench16.log
timeout 400 python -m pytest tests/test_ba_gpu.py tests/test_video_c5_gpu.py -m gpu -q -x > gpurun_out/r02_t16.log 2>&1; tail -4 gpurun_out/r02_t16.log
timeout 300 python tools/video_c5.py --final-only 2 > gpurun_out/r02_c5_final.log 2>&1; tail -2 gpurun_out/r02_c5_final.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c5_final.csv python tools/video_c5.py --final-only 1 > /dev/null 2>&1
ls -la gpurun_out/r02_launches_c5_final.csv
