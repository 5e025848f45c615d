set -x
timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "cholesky or c3" > gpurun_out/r02_chol_tests.log 2>&1; tail -3 gpurun_out/r02_chol_tests.log
timeout 120 python tools/microbench.py chol 2403 > gpurun_out/r02_chol_bench.log 2>&1
timeout 120 python tools/microbench.py ba >> gpurun_out/r02_chol_bench.log 2>&1
cat gpurun_out/r02_chol_bench.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
VGG_CORR_TC=0 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err
timeout 300 python -m pytest tests/test_corr_gpu.py tests/test_tracker_gpu.py -m gpu -q > gpurun_out/r02_corr_tests.log 2>&1; tail -15 gpurun_out/r02_corr_tests.log
timeout 300 python - > gpurun_out/r02_corr_tc.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.corr_section(torch.device('cuda:0'), 6540.5), indent=0))
PY
cat gpurun_out/r02_corr_tc.log
