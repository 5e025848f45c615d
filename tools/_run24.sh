set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:trsv_upper -s 5 -c 1 -o gpurun_out/r02_ncu_trsv -f python tools/microbench.py ba > gpurun_out/r02_ncu_trsv.log 2>&1
ls -la gpurun_out/r02_ncu_trsv.ncu-rep
