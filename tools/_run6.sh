set -x
timeout 400 python -m pytest tests/test_corr_gpu.py tests/test_ba_gpu.py tests/test_pycolmap_compat_gpu.py -m gpu -q -k "tensor_core or blocks or c3 or cholesky or compat or global_ba or pose_calls" > gpurun_out/r02_t6.log 2>&1; tail -6 gpurun_out/r02_t6.log
timeout 120 python tools/microbench.py chol 2403 > gpurun_out/r02_blocks_bench.log 2>&1
timeout 120 python tools/microbench.py blocks 4096 >> gpurun_out/r02_blocks_bench.log 2>&1
timeout 120 python tools/microbench.py blocks 131072 >> gpurun_out/r02_blocks_bench.log 2>&1
timeout 120 python tools/microbench.py ba >> gpurun_out/r02_blocks_bench.log 2>&1
cat gpurun_out/r02_blocks_bench.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
timeout 300 python - > gpurun_out/r02_corr_tc.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.corr_section(torch.device('cuda:0'), 6540.5)))
PY
cat gpurun_out/r02_corr_tc.log | cut -c1-1500
