#!/bin/bash
# r03a: 2^21 … 2^23 through the tile kernels: parity, then against the two-pass kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for e in "RONK_NTT3_MID=1" "RONK_NTT3_MID=0" "RONK_NTT3_MID=1 RONK_NTT3_T1=0"; do
  echo "$e $(env $e python tools/time_sizes.py 21:1 22:1 23:1 21:8 22:4 23:2 24:1 2>/dev/null)"
done
done
} 2>&1 | tee gpurun_out/sum/r03a_ab.txt
