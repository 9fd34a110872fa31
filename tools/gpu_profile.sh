#!/bin/bash
# Profiling pass on the GPU box (run through gpurun; everything lands in gpurun_out/<tag>_*):
#   tools/gpu_profile.sh <tag> [kernel-regex:skip:count:name ...]
# 1. the launch list of the default bench command (ncu --metrics gpu__time_duration.sum --clock-control none)
# 2. one `ncu --set full` capture per requested kernel, taken from an eager run of the same step (second step onwards,
#    so tables and caches are warm), summarised later with tools/ncu_summary.py
# Numbers printed by a run under ncu are never bench values.
set -u
TAG=${1:-prof}
shift || true
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
for spec in "$@"; do
    IFS=: read -r regex skip count name <<< "$spec"
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" --launch-skip "$skip" --launch-count "$count" \
        -o $OUT/${TAG}_ncu_${name} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph \
        > $OUT/${TAG}_ncu_${name}.log 2>&1
done
ls -la $OUT | grep "${TAG}_" || true
