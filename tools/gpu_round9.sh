#!/bin/bash
# r02l: ncu evidence for the instantiations the earlier captures lack: the inverse and fused-multiply
# 256-point-tile passes at 2^24, the config-5 shape (512 × 2^16), the 2^20 passes, the cluster kernel; launch list at HEAD.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
cap() {  # name, kernel regex, skip, count, workload...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/$name "$@" > gpurun_out/sum/$name.log 2>&1
  python tools/summarize_ncu.py gpurun_out/$name.ncu-rep gpurun_out/sum/$name 2>&1 | tail -1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02l_ntt3_24_variants "ntt3_kernel" 9 9 python tools/ncu_evidence.py ntt24
cap r02l_ntt3_16_config5 "ntt3_kernel" 4 4 python tools/ncu_evidence.py ntt16
cap r02l_ntt3_20 "ntt3_kernel|ntt3c_kernel" 6 6 python tools/ncu_evidence.py ntt20
cap r02l_ntt16_cluster "ntt16c_kernel" 2 2 python tools/ncu_evidence.py ntt16c
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/sum/r02l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -4 gpurun_out/sum/r02l_launches.csv | cut -c1-200
grep -E "==|time_duration|dram__bytes|inst_executed.sum|pipe_alu" gpurun_out/sum/r02l_*_metrics.txt | cut -c1-220
