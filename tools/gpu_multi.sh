#!/bin/bash
# multi-GPU validation + the driver's own N-GPU bench command (run under `gpurun --gpus N`): bash tools/gpu_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>gpurun_out/multi_check_${N}.err | tee gpurun_out/r02_multi_gpu_check_${N}x.json | cut -c1-1500
tail -5 gpurun_out/multi_check_${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_${N}.err | tee gpurun_out/r02_bench_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['copy_gbs_per_rank'],d['e2e']['numa']); print(json.dumps(d.get('multi'),indent=0)[:3500])"
tail -5 gpurun_out/bench_${N}.err
