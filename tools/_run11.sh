set -x
timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "cholesky" > gpurun_out/r02_t11.log 2>&1; tail -3 gpurun_out/r02_t11.log
for f in 1 0; do
  VGG_CHOL_FUSE=$f timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench11.log 2>&1
  VGG_CHOL_FUSE=$f timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench11.log 2>&1
done
VGG_CHOL_GRAPH=0 timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench11.log 2>&1
cat gpurun_out/r02_bench11.log
timeout 400 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "c3 or global_ba or c2" >> gpurun_out/r02_t11.log 2>&1; tail -3 gpurun_out/r02_t11.log
