#!/bin/bash
# Multi-GPU measurements on one box (run through gpurun --gpus N): tools/multi_gpu_round.sh <tag> <N> [c5_spp]
#   training step (C2, one view per rank), row-band renders C5 (spp c5_spp, default the configuration's 128) and C3 (spp 512),
#   and the bit-identity check of the banded render against one GPU.  Every torchrun is under its own timeout.
TAG=$1; N=$2; C5SPP=${3:-128}
OUT=gpurun_out
mkdir -p $OUT
run() { # name, timeout, args...
  local name=$1 t=$2; shift 2
  if [ "$N" = "1" ]; then timeout $t python "$@" > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  else timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@" > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err; fi
  echo "$name rc=$?"
  python - $OUT/${TAG}_${name}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("   N=%d  %.3f ms/step  e2e %.3f ms  value %.4g  %s  %s" % (d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"],
          d.get("bands", {}).get("bounds", ""), d.get("per_rank_ms_per_step", {}).get("device", "")))
    print("   ", d.get("execution", "")[:110], "|", d.get("collective"))
except Exception as e:
    print("   no JSON line:", e)
PY
}
run train 240 bench.py --gpus $N --steps 20 --no-cpu-baseline
run c5 400 bench.py --gpus $N --config C5 --spp $C5SPP --steps 3 --no-cpu-baseline
run c3 400 bench.py --gpus $N --config C3 --steps 3 --no-cpu-baseline
if [ "$N" != "1" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + RANDOM % 90)) tools/check_rows_sharded.py 6 C5 2>&1 | grep -E "row bands|Error|error" | tail -3
fi
