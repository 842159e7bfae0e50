#!/bin/bash
# r02o: A/B of the two-chain stepping of the pass-1 twiddle against HEAD (5 CTAs per SM, FMA_TAIL 3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_X=base"
run "RONK_LIB_PATH=$PWD/variants/libronk_step2.so"
run "RONK_X=base"
run "RONK_LIB_PATH=$PWD/variants/libronk_step2.so"
echo "base $(python tools/time_sizes.py 20:1 20:4 16:1 16:512 19:1 22:1 2>/dev/null)"
echo "step2 $(RONK_LIB_PATH=$PWD/variants/libronk_step2.so python tools/time_sizes.py 20:1 20:4 16:1 16:512 2>/dev/null)"
} 2>&1 | tee gpurun_out/sum/r02o_ab.txt
