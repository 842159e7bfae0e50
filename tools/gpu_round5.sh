#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py tests/test_gpu_kzg.py -m gpu -x -q 2>&1 | tail -3
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_LIB_PATH=$PWD/variants/libronk_unr2.so"
run "RONK_NTT3_T1=0"
run "RONK_NTT3_T1=1"
run "RONK_LIB_PATH=$PWD/variants/libronk_unr2.so RONK_NTT3_T1=1"
run "RONK_NTT3_T1=0"
python tests/config_timing.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['single_transform_ms']); print({k:(v.get('ms') or v.get('call_ms')) for k,v in d.items() if k!='single_transform_ms'}); print(d.get('config4_msm_2^20'))"
