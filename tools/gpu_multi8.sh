#!/bin/bash
# r02k: 8-GPU validation (one box): the coord commit kernel on GPU 0, the C-ABI multi-GPU modes bit-exact at G = 8 and
# G = 4 (NCCL and fused peer-memory flavours), then the driver's own 8-GPU bench command with its `multi` block.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
python -m pytest tests/test_gpu_kzg.py -m gpu -x -q 2>&1 | tail -2
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/multi_gpu_check.py 2>gpurun_out/sum/multi_check_${N}.err | tee gpurun_out/sum/r02k_multi_gpu_check_${N}x.json | cut -c1-1800
  tail -3 gpurun_out/sum/multi_check_${N}.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/sum/bench_8.err | tee gpurun_out/sum/r02k_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['copy_gbs_per_rank'],d['e2e']['numa']); print(json.dumps(d.get('multi'),indent=0)[:3500])"
tail -3 gpurun_out/sum/bench_8.err
nvidia-smi topo -m > gpurun_out/sum/r02k_topo.txt 2>&1
