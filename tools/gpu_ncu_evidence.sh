#!/bin/bash
# ncu captures for profiles/r02_* (1 GPU).  Kept short: --set full replays every kernel ~40 times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ntt_tile_kernel" -s 6 -c 6 -o gpurun_out/r02g_ntt24_variants python tools/ncu_evidence.py ntt24 > gpurun_out/ncu_g1.log 2>&1; tail -1 gpurun_out/ncu_g1.log
ncu --set full --clock-control none --import-source on -k regex:"ntt_tile_kernel" -s 4 -c 4 -o gpurun_out/r02g_ntt16_config5 python tools/ncu_evidence.py ntt16 > gpurun_out/ncu_g2.log 2>&1; tail -1 gpurun_out/ncu_g2.log
ncu --set full --clock-control none --import-source on -k regex:"binop_kernel|div_linear" -s 0 -c 8 -o gpurun_out/r02g_field python tools/ncu_evidence.py field > gpurun_out/ncu_g3.log 2>&1; tail -1 gpurun_out/ncu_g3.log
ncu --set full --clock-control none --import-source on -k regex:"msm_hist" -s 2 -c 2 -o gpurun_out/r02g_msm_2_20 python tools/ncu_evidence.py msm > gpurun_out/ncu_g4.log 2>&1; tail -1 gpurun_out/ncu_g4.log
# the launch list of the default bench command (shares of the step, cold-cache serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_g5.log 2>&1; tail -1 gpurun_out/ncu_g5.log
ncu --set full --clock-control none --import-source on -k regex:"ntt_tile_kernel" -s 8 -c 2 -o gpurun_out/r02g_ntt_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_g6.log 2>&1; tail -1 gpurun_out/ncu_g6.log
python bench.py > gpurun_out/r02g_bench_default.json 2>gpurun_out/bench_g.err; tail -2 gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_default.json')); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_ntt']['frac'], d['spot_check'], d['clocks'])"
