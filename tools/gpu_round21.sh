#!/bin/bash
# r03e: split transforms (2^17 … 2^19 batched, 2^25, 2^26): parity, then against the two-pass kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
python -m pytest tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -3
for e in "RONK_NTT3_SPLIT=1" "RONK_NTT3_SPLIT=0" "RONK_NTT3_SPLIT=1 RONK_NTT3_SPLIT_MIN16=1"; do
  echo "$e $(env $e python tools/time_sizes.py 25:1 26:1 17:128 18:64 19:32 17:8 18:4 19:2 17:1 18:1 19:1 2>/dev/null)"
done
} 2>&1 | tee gpurun_out/sum/r03e_ab.txt
