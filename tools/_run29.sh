set -x
timeout 300 python -m pytest tests/test_video_c5_gpu.py tests/test_video_gpu.py -m gpu -q -x > gpurun_out/r02_t29.log 2>&1; tail -3 gpurun_out/r02_t29.log
timeout 300 python tools/video_c5.py --final-only 3 > gpurun_out/r02_c5_final_band4.log 2>&1; tail -1 gpurun_out/r02_c5_final_band4.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ba_blocks -c 6 --csv --log-file gpurun_out/r02_launches_c5_blocks.csv python tools/video_c5.py --final-only 1 > /dev/null 2>&1; tail -2 gpurun_out/r02_launches_c5_blocks.csv | cut -c150-330
