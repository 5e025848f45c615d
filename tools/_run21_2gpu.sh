set -x
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x > gpurun_out/r02_dist_tests2.log 2>&1; tail -3 gpurun_out/r02_dist_tests2.log
bash tools/_run_ngpu.sh 2
