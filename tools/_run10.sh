set -x
timeout 120 python tools/microbench.py chol128 > gpurun_out/r02_bench10.log 2>&1
timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "cholesky" > gpurun_out/r02_t10.log 2>&1; tail -4 gpurun_out/r02_t10.log
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  VGG_CHOL_LEAF=$1 VGG_CHOL_FUSE=$2 timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench10.log 2>&1
  VGG_CHOL_LEAF=$1 VGG_CHOL_FUSE=$2 timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench10.log 2>&1
done
cat gpurun_out/r02_bench10.log
timeout 400 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "c3 or cholesky or global_ba or c2" >> gpurun_out/r02_t10.log 2>&1; tail -4 gpurun_out/r02_t10.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol3.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
