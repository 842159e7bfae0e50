#!/bin/bash
# r02r: run-time switches at HEAD (no rebuild): pass-1 twiddle table, PDL, one-group-per-thread threshold
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_X=base"
run "RONK_NTT3_T1=1"
run "RONK_NTT3_PDL=0"
run "RONK_X=base"
for e in "RONK_X=base" "RONK_NTT3_NG1_TILES=0" "RONK_NTT3_NG1_TILES=6" "RONK_NTT3_NG1_TILES=12" "RONK_NTT3_PDL=0"; do
  echo "$e $(env $e python tools/time_sizes.py 20:1 20:2 20:4 16:1 16:4 16:16 16:64 16:128 2>/dev/null)"
done
} 2>&1 | tee gpurun_out/sum/r02r_switches.txt
