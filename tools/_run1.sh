set -x
python -c "import pycolmap" > gpurun_out/r02_probe_pycolmap.txt 2>&1; python -c "import pyceres" >> gpurun_out/r02_probe_pycolmap.txt 2>&1
(pip download pycolmap==3.10.0 --no-deps -d /tmp/x 2>&1 | tail -3) >> gpurun_out/r02_probe_pycolmap.txt; nproc >> gpurun_out/r02_probe_pycolmap.txt
timeout 300 python -m pytest tests/test_ba_gpu.py tests/test_syrk_i8_gpu.py -m gpu -q -k "cholesky or syrk or float64 or nonfinite" > gpurun_out/r02_chol_tests.log 2>&1; tail -15 gpurun_out/r02_chol_tests.log
for v in "" "VGG_CHOL_GRAPH=0" "VGG_CHOL_GRAPH=0 VGG_CHOL_LOOKAHEAD=0" "VGG_CHOL_LOOKAHEAD=0"; do env $v timeout 120 python tools/microbench.py chol 2403; done > gpurun_out/r02_chol_bench.log 2>&1
for v in "" "VGG_CHOL=lib"; do env $v timeout 120 python tools/microbench.py ba; done >> gpurun_out/r02_chol_bench.log 2>&1
cat gpurun_out/r02_chol_bench.log
timeout 600 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_gputest_a.log 2>&1; tail -25 gpurun_out/r02_gputest_a.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_chol.csv python tools/microbench.py chol 2403 > /dev/null 2>&1
