#!/bin/bash
# A/B harness used throughout the optimisation log: build the library several times with
# different -D switches into variants/ (git-ignored, but it travels to the GPU box), then time each one with
# bench.py on the GPU in ONE gpurun call.
#
#   tools/ab_variants.sh build  base ""  t19 "-DRONK_FMA_TAIL=19"  lb16 "-DRONK_LD_BATCH=16"
#   gpurun --timeout 600 -- 'bash tools/ab_variants.sh run base t19 lb16 base'
#
# `run` prints one line per variant: name, ms per 2^24 transform, per-kernel ms.  Repeat the first name last to
# see the run-to-run noise (≈ ±0.3 %).  The in-tree library is restored by a final plain `make`.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
case "$1" in
  build)
    shift
    mkdir -p "$ROOT/variants"
    while [ $# -ge 2 ]; do
      name="$1"; flags="$2"; shift 2
      make -B -C "$ROOT/ronkathon_b200/csrc" OUT="$ROOT/variants/libronk_$name.so" EXTRA="$flags" 2>&1 | grep -E " error" && exit 1
      echo "built variants/libronk_$name.so  [$flags]"
    done
    make -B -C "$ROOT/ronkathon_b200/csrc" 2>&1 | grep -E " error" && exit 1
    echo "in-tree library rebuilt with the default switches"
    ;;
  run)
    shift
    for name in "$@"; do
      RONK_LIB_PATH="$ROOT/variants/libronk_$name.so" python "$ROOT/bench.py" --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null |
        python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('$name', round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['roofline']['kernel_ms'].items()})"
    done
    ;;
  *)
    sed -n 2,12p "$0"; exit 2 ;;
esac
