set -x
timeout 120 python tools/microbench.py trsv 2402 > gpurun_out/r02_trsv_probe3.log 2>&1; head -6 gpurun_out/r02_trsv_probe3.log | cut -c1-400; tail -2 gpurun_out/r02_trsv_probe3.log
