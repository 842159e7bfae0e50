#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/one_ntt24.py 2>&1 | tail -5
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py -m gpu -x -q 2>&1 | tail -3
for v in "RONK_NTT3=0" "RONK_NTT3=1 RONK_NTT3_PDL=0" "RONK_NTT3=1 RONK_NTT3_PDL=1" "RONK_NTT3=1 RONK_NTT3_PDL=0" "RONK_NTT3=1 RONK_NTT3_PDL=1"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"
done
ncu --set full --clock-control none --import-source on -k regex:"ntt3_kernel" -s 9 -c 3 -o gpurun_out/r02h_ntt3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sum/r02h_ntt3.log 2>&1
python tools/summarize_ncu.py gpurun_out/r02h_ntt3.ncu-rep gpurun_out/sum/r02h_ntt3 2>&1 | tail -1
rm -f gpurun_out/r02h_ntt3.ncu-rep
cat gpurun_out/sum/r02h_ntt3_metrics.txt | grep -E "==|time_duration|inst_executed.sum|pipe_alu|issue_active|bank_conflicts|registers|dram__bytes|stall cycles|warps_active"
