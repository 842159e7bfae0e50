#!/bin/bash
# r02k: the 2^20 interleaved-tile path + the polished commit kernel: parity tests, config timing, then the A/B of the
# arithmetic switches on the 256-point-tile transform (variants/ built by tools/build_ntt_variant.sh).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_kzg.py tests/test_gpu_poly.py -m gpu -x -q 2>&1 | tail -4
python tests/config_timing.py 2>/dev/null > gpurun_out/sum/r02k_config_timing.json; cut -c1-2600 gpurun_out/sum/r02k_config_timing.json
RONK_NTT3_20=0 python tests/config_timing.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('RONK_NTT3_20=0', d['config2_ntt_2^20'])"
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_X=base"
for v in t0 t7 t11 t19 t31 unr2 lb5; do run "RONK_LIB_PATH=$PWD/variants/libronk_$v.so"; done
run "RONK_X=base"
