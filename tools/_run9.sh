set -x
timeout 120 python tools/microbench.py chol128 > gpurun_out/r02_bench9.log 2>&1
for leaf in 0 1; do
  VGG_CHOL_LEAF=$leaf timeout 120 python tools/microbench.py chol 2403 >> gpurun_out/r02_bench9.log 2>&1
  VGG_CHOL_LEAF=$leaf timeout 120 python tools/microbench.py ba >> gpurun_out/r02_bench9.log 2>&1
done
cat gpurun_out/r02_bench9.log
VGG_CHOL_LEAF=1 timeout 400 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "c3 or cholesky or global_ba" > gpurun_out/r02_t9.log 2>&1; tail -4 gpurun_out/r02_t9.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:corr_tc_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_corr_tc8 -f python tools/profile_r02.py corr_tc > gpurun_out/r02_ncu_b.log 2>&1
timeout 300 $NCU -k regex:ba_blocks -s 2 -c 1 -o gpurun_out/r02_ncu_blocks_c3b -f python tools/profile_r02.py blocks > gpurun_out/r02_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
