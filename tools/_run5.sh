set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:chol_panel -s 19 -c 2 -o gpurun_out/r02_ncu_chol_panel -f python tools/profile_r02.py chol > gpurun_out/r02_ncu_a.log 2>&1
timeout 300 $NCU -k regex:corr_tc_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_corr_tc -f python tools/profile_r02.py corr_tc > gpurun_out/r02_ncu_b.log 2>&1
timeout 300 $NCU -k regex:corr_sample -s 1 -c 1 -o gpurun_out/r02_ncu_corr_fine -f python tools/profile_r02.py corr_fine > gpurun_out/r02_ncu_c.log 2>&1
timeout 300 $NCU -k regex:ba_blocks -s 2 -c 1 -o gpurun_out/r02_ncu_blocks_c3 -f python tools/profile_r02.py blocks > gpurun_out/r02_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_a.log gpurun_out/r02_ncu_b.log
