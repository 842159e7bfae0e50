#!/bin/bash
# r02w: the 8 MiB pass-A2 twiddle table of the 2^20-point transform: parity, then with / without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
python -m pytest tests/test_gpu_ntt.py -m gpu -x -q -k "2_20 or opt_in" 2>&1 | tail -2
for rep in 1 2; do
for e in "RONK_NTT3_T1=1" "RONK_NTT3_T1=0"; do
  echo "$e $(env $e python tools/time_sizes.py 20:1 20:2 20:4 20:16 24:1 2>/dev/null)"
done
done
} 2>&1 | tee gpurun_out/sum/r02x_ab.txt
