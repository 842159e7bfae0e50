#!/bin/bash
# One GPU call's worth of checks used while tuning (run under gpurun from the repo root).
cd "$(dirname "$0")/.."
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/one_ntt24.py 2>&1 | tail -8
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py tests/test_gpu_kzg.py tests/test_gpu_cpp_mirror.py -m gpu -x -q 2>&1 | tail -3
for v in "RONK_FAST12=0" "RONK_FAST12=1" "RONK_FAST12=1"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"
done
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/r02f_bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']); print(json.dumps(d['configs'], indent=0)[:2500])"
tail -3 gpurun_out/bench_err.log
ncu --set full --clock-control none --import-source on -k regex:ntt12 -s 6 -c 2 -o gpurun_out/r02f_ntt12 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log
