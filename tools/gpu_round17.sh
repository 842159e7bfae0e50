#!/bin/bash
# r02t: prefetch of the pass-1 twiddle-table rows before the network (L1 / L2) against HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
for rep in 1 2; do
run "RONK_X=base"
for v in e1 e3 e5 e7; do run "RONK_LIB_PATH=$PWD/variants/libronk_$v.so"; done
done
echo "base $(python tools/time_sizes.py 20:1 16:512 16:1 2>/dev/null)"
echo "e7 $(RONK_LIB_PATH=$PWD/variants/libronk_e7.so python tools/time_sizes.py 20:1 16:512 16:1 2>/dev/null)"
} 2>&1 | tee gpurun_out/sum/r02u_ab.txt
