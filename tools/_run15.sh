set -x
timeout 300 python tools/video_c5.py --frames 160 --new 128 -v > gpurun_out/r02_c5_small.log 2>&1; tail -12 gpurun_out/r02_c5_small.log | cut -c1-900
timeout 900 python tools/video_c5.py --frames 1000 --new 512 -v --json gpurun_out/r02_c5.json > gpurun_out/r02_c5.log 2>&1; tail -16 gpurun_out/r02_c5.log | cut -c1-900
