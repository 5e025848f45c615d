set -x
timeout 120 python tools/microbench.py trsv 2402 > gpurun_out/r02_trsv_probe.log 2>&1; cat gpurun_out/r02_trsv_probe.log | head -50
