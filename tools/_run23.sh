set -x
rm -f gpurun_out/r02_bench23.log
for m in flag data; do
  VGG_TRSV_POLL=$m timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench23.log 2>&1
done
cat gpurun_out/r02_bench23.log
timeout 400 python -m pytest tests/test_ba_gpu.py tests/test_pipeline_gpu.py tests/test_video_c5_gpu.py -m gpu -q -x > gpurun_out/r02_t23.log 2>&1; tail -3 gpurun_out/r02_t23.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:trsv_upper -c 30 --csv --log-file gpurun_out/r02_launches_trsv_flag.csv python tools/microbench.py ba > /dev/null 2>&1
VGG_TRSV_POLL=data timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:trsv_upper -c 30 --csv --log-file gpurun_out/r02_launches_trsv_data.csv python tools/microbench.py ba > /dev/null 2>&1
tail -2 gpurun_out/r02_launches_trsv_flag.csv | cut -c1-300; tail -2 gpurun_out/r02_launches_trsv_data.csv | cut -c1-300
