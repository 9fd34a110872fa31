#!/bin/bash
# A/B of launch-shape tuning on the GPU box: tools/bench_variants.sh <tag> "<bench args 1>" "<bench args 2>" ...
# prints ms/step (device, e2e) and the heaviest entry points of every variant; full JSON lines in gpurun_out/<tag>_<k>.json
TAG=$1; shift
k=0
for v in "$@"; do
  python bench.py --steps 10 --no-cpu-baseline $v > gpurun_out/${TAG}_$k.json 2> gpurun_out/${TAG}_$k.err
  python - "$v" gpurun_out/${TAG}_$k.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("%-44s %.3f ms  e2e %.3f ms  " % (sys.argv[1] or "(default)", d["ms_per_step"], d["e2e"]["ms_per_step"]),
          {k: v for k, v in d["kernel_ms_per_step"].items() if v > 0.3})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  k=$((k+1))
done
