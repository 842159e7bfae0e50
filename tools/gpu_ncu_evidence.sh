#!/bin/bash
# ncu captures for profiles/r02g_* (1 GPU).  --set full replays every kernel ~40 times: keep the counts small.  The
# reports are summarised ON the box (tools/summarize_ncu.py) and deleted: gpurun brings back at most 64 MiB.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
cap() {  # name, kernel regex, skip, count, workload...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/$name "$@" > gpurun_out/sum/$name.log 2>&1
  python tools/summarize_ncu.py gpurun_out/$name.ncu-rep gpurun_out/sum/$name 2>&1 | tail -1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02g_ntt_final "ntt_tile_kernel" 8 2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
cap r02g_ntt24_variants "ntt_tile_kernel" 8 4 python tools/ncu_evidence.py ntt24
cap r02g_ntt16_config5 "ntt_tile_kernel" 4 4 python tools/ncu_evidence.py ntt16
cap r02g_field "binop_kernel|div_linear" 0 8 python tools/ncu_evidence.py field
cap r02g_msm_2_20 "msm_hist" 2 2 python tools/ncu_evidence.py msm
# the launch list of the default bench command (shares of the step, cold-cache serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/sum/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -4 gpurun_out/sum/r02g_launches.csv | cut -c1-200
