set -x
timeout 600 python -m pytest tests/test_video_c5_gpu.py tests/test_ba_gpu.py tests/test_video_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r02_t20.log 2>&1; tail -4 gpurun_out/r02_t20.log
timeout 300 python tools/_diag_band.py > gpurun_out/r02_diag_band3.log 2>&1; tail -12 gpurun_out/r02_diag_band3.log
timeout 300 python tools/video_c5.py --final-only 3 > gpurun_out/r02_c5_final_band3.log 2>&1; tail -1 gpurun_out/r02_c5_final_band3.log | cut -c1-400
timeout 900 python tools/video_c5.py --frames 1000 --new 512 --json gpurun_out/r02_c5_band3.json > gpurun_out/r02_c5_band3.log 2>&1; tail -1 gpurun_out/r02_c5_band3.log | cut -c1-1200
timeout 120 python tools/microbench.py ba > gpurun_out/r02_bench20.log 2>&1; cat gpurun_out/r02_bench20.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c5_final_band3.csv python tools/video_c5.py --final-only 1 > /dev/null 2>&1
