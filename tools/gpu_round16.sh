#!/bin/bash
# r03f: HEAD evidence — every GPU test, compute-sanitizer on the smoke workload, the default bench line, per-config
# timing, the launch list and one `ncu --set full` capture of the three 256-point-tile passes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/sum/r03f_gpu_tests.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_smoke.py 2>&1 | tail -4 | tee gpurun_out/sum/r03f_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_smoke.py 2>&1 | tail -4 | tee gpurun_out/sum/r03f_sanitizer_racecheck.txt
python bench.py --steps 20 --warmup 5 2>gpurun_out/sum/bench_err.log | tee gpurun_out/sum/r03f_bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], d['roofline']['whole_ntt'], 'traffic', d['roofline']['traffic']); print('e2e', d['e2e']['ms_per_step']); print('spot', d['spot_check']['ok']); print(json.dumps(d['configs'])[:1500]); print(d['clocks'])"
tail -3 gpurun_out/sum/bench_err.log
python tests/config_timing.py 2>/dev/null > gpurun_out/sum/r03f_config_timing.json; cut -c1-2600 gpurun_out/sum/r03f_config_timing.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/sum/r03f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -3 gpurun_out/sum/r03f_launches.csv | cut -c1-220
ncu --set full --clock-control none --import-source on -k regex:"ntt3_kernel" -s 9 -c 3 -o gpurun_out/r03f_ntt3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sum/r03f_ntt3.log 2>&1
python tools/summarize_ncu.py gpurun_out/r03f_ntt3.ncu-rep gpurun_out/sum/r03f_ntt3 2>&1 | tail -1
rm -f gpurun_out/r03f_ntt3.ncu-rep
grep -E "==|time_duration|inst_executed.sum|pipe_alu|pipe_fma|issue_active|registers|dram__bytes|stall cycles|warps_active" gpurun_out/sum/r03f_ntt3_metrics.txt | cut -c1-230
