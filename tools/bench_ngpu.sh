# usage: bash tools/bench_ngpu.sh N   (inside gpurun --gpus N)
N=$1
set -x
nvidia-smi -L | head -8
for v in 2 0; do
  VGG_FABRIC=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus $N --steps 5 --warmup 3 --no-corr --no-c5 > gpurun_out/r02_bench_${N}gpu_fabric$v.json 2> gpurun_out/r02_bench_${N}gpu_fabric$v.err
  grep '^{' gpurun_out/r02_bench_${N}gpu_fabric$v.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=$N fabric=$v', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'cost', d['config'].get('final_cost'), 'nccl_nranks', d['config'].get('nccl_nranks'))
"
  tail -2 gpurun_out/r02_bench_${N}gpu_fabric$v.err
done
