#!/bin/bash
# r02m: second A/B round on the 256-point-tile transform (register budget × tail switches), batched 2^20 with and
# without the interleaved-tile path, cluster kernel switch-over.  Everything goes to gpurun_out/sum/r02m_ab.txt.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_X=base"
for v in lb5 lb5t19 lb5t11 lb4 lb4t19 t19; do run "RONK_LIB_PATH=$PWD/variants/libronk_$v.so"; done
run "RONK_X=base"
echo "--- batched sizes (ms per call)"
for e in "RONK_NTT3_20=1" "RONK_NTT3_20=0" "RONK_LIB_PATH=$PWD/variants/libronk_lb5.so" "RONK_LIB_PATH=$PWD/variants/libronk_lb5t19.so"; do
  echo "$e $(env $e python tools/time_sizes.py 20:1 20:2 20:4 20:16 16:1 16:8 16:9 16:64 16:512 2>/dev/null)"
done
echo "RONK_NTT16_CLUSTER_MAX_BATCH=0 $(RONK_NTT16_CLUSTER_MAX_BATCH=0 python tools/time_sizes.py 16:1 16:2 16:4 16:8 2>/dev/null)"
echo "RONK_NTT16_CLUSTER_MAX_BATCH=64 $(RONK_NTT16_CLUSTER_MAX_BATCH=64 python tools/time_sizes.py 16:1 16:2 16:4 16:8 16:16 16:32 16:64 2>/dev/null)"
} 2>&1 | tee gpurun_out/sum/r02m_ab.txt
