#!/bin/bash
# Build variants/libronk_<name>.so that differs from the in-tree library only in ntt.cu's -D switches:
#   tools/build_ntt_variant.sh <name> "<flags>"      (the other objects are the in-tree ones: run `make` first)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
C="$ROOT/ronkathon_b200/csrc"
mkdir -p "$ROOT/variants"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c "$C/ntt.cu" -o "$ROOT/variants/ntt_$1.o"
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/variants/libronk_$1.so" "$ROOT/variants/ntt_$1.o" "$C/api.o" "$C/field_ops.o" "$C/poly.o" "$C/msm.o" "$C/dist.o" -cudart static -ldl
rm -f "$ROOT/variants/ntt_$1.o"
echo "built variants/libronk_$1.so [$2]"
