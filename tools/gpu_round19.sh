#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sum
{
for rep in 1 2; do
echo "base $(python tools/time_polymul.py 2>/dev/null)"
echo "e9 $(RONK_LIB_PATH=$PWD/variants/libronk_e9.so python tools/time_polymul.py 2>/dev/null)"
done
} 2>&1 | tee gpurun_out/sum/r02y_ab.txt
