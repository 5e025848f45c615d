set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_t28.log 2>&1; tail -4 gpurun_out/r02_t28.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_final2.json 2> gpurun_out/r02_bench_1gpu_final2.err; tail -c 1500 gpurun_out/r02_bench_1gpu_final2.json; tail -2 gpurun_out/r02_bench_1gpu_final2.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
