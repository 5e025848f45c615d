set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_t22.log 2>&1; tail -5 gpurun_out/r02_t22.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; tail -c 3000 gpurun_out/r02_bench_1gpu_final.json; tail -3 gpurun_out/r02_bench_1gpu_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/r02_bench_reference_arm_final.err; tail -c 1200 gpurun_out/r02_bench_reference_arm_final.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-corr --no-c5 > gpurun_out/r02_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:chol_panel -s 19 -c 2 -o gpurun_out/r02_ncu_chol_panel2 -f python tools/profile_r02.py chol > gpurun_out/r02_ncu_a2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
