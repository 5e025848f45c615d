set -x
timeout 120 python tools/microbench.py chol128 > gpurun_out/r02_bench12.log 2>&1
timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench12.log 2>&1
VGG_CHOL_TIMING_SKIP_BULK=1 timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench12.log 2>&1
timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench12.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv >> gpurun_out/r02_bench12.log
cat gpurun_out/r02_bench12.log
timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "cholesky or c3" > gpurun_out/r02_t12.log 2>&1; tail -3 gpurun_out/r02_t12.log
