#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "RONK_PDL=0" "RONK_PDL=1" "RONK_PDL=0" "RONK_PDL=1"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"
done
for v in "RONK_PDL=0" "RONK_PDL=1"; do env $v python tests/config_timing.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['single_transform_ms']); print({k:(v.get('ms') or v.get('call_ms')) for k,v in d.items() if k!='single_transform_ms'})"; done
bash tools/gpu_ncu_evidence.sh
python bench.py > gpurun_out/sum/r02g_bench_default.json 2>gpurun_out/sum/bench_g.err; tail -2 gpurun_out/sum/bench_g.err
python tests/config_timing.py > gpurun_out/sum/r02g_config_timing.json 2>/dev/null
du -sh gpurun_out
