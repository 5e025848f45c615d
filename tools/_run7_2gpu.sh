set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x > gpurun_out/r02_dist_tests.log 2>&1; tail -15 gpurun_out/r02_dist_tests.log
for v in 2 1 0; do
  VGG_FABRIC=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-corr > gpurun_out/r02_bench_2gpu_fabric$v.json 2> gpurun_out/r02_bench_2gpu_fabric$v.err
  tail -c 1200 gpurun_out/r02_bench_2gpu_fabric$v.json; tail -3 gpurun_out/r02_bench_2gpu_fabric$v.err
done
