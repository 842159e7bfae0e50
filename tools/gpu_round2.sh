#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -2
for v in "RONK_FAST12=0" "RONK_FAST12=1 RONK_PF_DIST2=0" "RONK_FAST12=1 RONK_PF_DIST2=1" "RONK_FAST12=1 RONK_PF_DIST2=2" "RONK_FAST12=1 RONK_PF_DIST2=1 RONK_PF_DIST=2"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"
done
