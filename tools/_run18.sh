set -x
timeout 300 python -m pytest tests/test_syrk_i8_gpu.py -m gpu -q -x > gpurun_out/r02_t18.log 2>&1; tail -5 gpurun_out/r02_t18.log
timeout 600 python tools/_diag_band.py > gpurun_out/r02_diag_band.log 2>&1; tail -14 gpurun_out/r02_diag_band.log
