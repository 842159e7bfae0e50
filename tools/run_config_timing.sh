#!/bin/bash
# GPU box helper: config timings with the in-tree library and, if present, a comparison build in variants/
python tests/config_timing.py > gpurun_out/r01j_config_timing.json 2>gpurun_out/r01j_ct.err; tail -2 gpurun_out/r01j_ct.err
if [ -n "$RONK_AB_ENV" ]; then   # A/B: the same library with one experiment switch set, e.g. RONK_AB_ENV="RONK_TILE_ADAPT=0"
  env $RONK_AB_ENV python tests/config_timing.py > gpurun_out/r01j_config_timing_before.json 2>/dev/null
fi
python - <<'PY'
import json, os
for f in ("r01j_config_timing.json", "r01j_config_timing_before.json"):
    if not os.path.exists("gpurun_out/" + f): continue
    d = json.load(open("gpurun_out/" + f)); print(f)
    for k, v in d.items():
        print("  ", k, {kk: vv for kk, vv in v.items() if kk not in ("point_adds_per_s", "field_muls_per_s")} if "ms" in v or "call_ms" in v else v)
PY
