#!/bin/bash
# r02j: kzg::commit in group coordinates (msm_coord_kernel) — parity, timing against the histogram path, ncu; and the
# A/B of the pass-1 twiddle table of the 256-point-tile transform.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_plonk_polys.py tests/test_gpu_cpp_mirror.py -m gpu -x -q 2>&1 | tail -4
for v in "RONK_MSM_COORD=1" "RONK_MSM_COORD=0"; do
  env $v python tests/config_timing.py --only msm 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', {k:(v.get('call_ms'), v.get('kernel_ms'), v.get('bit_exact')) for k,v in d.items() if 'msm' in k})"
done
run() { env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'spot', d['spot_check']['ok'])"; }
run "RONK_NTT3_T1=0"
run "RONK_NTT3_T1=1"
run "RONK_NTT3_T1=0"
run "RONK_NTT3_T1=1"
ncu --set full --clock-control none --import-source on -k regex:"msm_coord" -s 2 -c 2 -o gpurun_out/r02j_msm_coord python tools/ncu_evidence.py msm > gpurun_out/sum/r02j_msm.log 2>&1
python tools/summarize_ncu.py gpurun_out/r02j_msm_coord.ncu-rep gpurun_out/sum/r02j_msm_coord_2_20 2>&1 | tail -1
rm -f gpurun_out/r02j_msm_coord.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"msm_coord" -s 2 -c 1 -o gpurun_out/r02j_msm_coord24 python tools/ncu_evidence.py msm24 > gpurun_out/sum/r02j_msm24.log 2>&1
python tools/summarize_ncu.py gpurun_out/r02j_msm_coord24.ncu-rep gpurun_out/sum/r02j_msm_coord_2_24 2>&1 | tail -1
rm -f gpurun_out/r02j_msm_coord24.ncu-rep
grep -E "==|time_duration|dram__bytes|dram_throughput|inst_executed.sum|issue_active|bank_conflicts|registers|stall" gpurun_out/sum/r02j_msm_coord_2_20_metrics.txt gpurun_out/sum/r02j_msm_coord_2_24_metrics.txt | cut -c1-200
