#!/bin/bash
# End-of-round evidence on ONE B200 (run through gpurun): ncu captures of the kernels VERDICT r1 names at the final state,
# the DRAM traffic of the dominant kernel (-> profiles/ncu_traffic_C2.json, read by bench.py), then tests, the default
# bench line, the reference arm and the single-GPU render lines.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-final}
COMMIT=${2:-unknown}   # the commit of the snapshot (the box has no .git): tools/final_round.sh <tag> $(git rev-parse --short HEAD)
OUT=gpurun_out
mkdir -p $OUT
tools/gpu_profile.sh $TAG "k_trace_any_persistent:12:3:any" "k_trace_closest_persistent:5:2:closest" "k_trace_mixed_persistent:8:2:mixed" \
    "initial_gen_px:4:1:initial_gen" "bounce_shade_gen_px:8:2:bounce_shade_gen" "k_hierarchy:1:1:hierarchy" "k_sort_pass:3:3:sort_pass" \
    "k_leaves:1:1:leaves" > $OUT/${TAG}_profile.log 2>&1
for k in any closest mixed initial_gen bounce_shade_gen hierarchy sort_pass leaves; do
  python tools/ncu_summary.py $OUT/${TAG}_ncu_$k.ncu-rep $OUT/${TAG}_ncu_$k.txt > /dev/null 2>&1
done
python - $TAG $COMMIT <<'PY'
import json, re, sys
tag, commit = sys.argv[1], sys.argv[2]
txt = open("gpurun_out/%s_ncu_any.txt" % tag).read().split("kernel ")
# kernel 1 of the capture = the boolean-ray launch of the spatial pass (launches 12, 13, 14 of the step: initial, spatial, visibility)
blk = [b for b in txt if b.startswith("1:")][0]
def val(name):
    m = re.search(name + r"\s+([0-9.,]+)\s+(\S+)", blk)
    v = float(m.group(1).replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump({"entry_point": "spatial_resampling", "kernel": "mr::k_trace_any_persistent<false> (boolean rays of the spatial pass)",
           "dram_bytes_per_launch": traffic, "source": "profiles/%s_ncu_any.txt (kernel 1; ncu --set full --clock-control none, eager run of bench.py)" % tag,
           "commit": commit},
          open("profiles/ncu_traffic_C2.json", "w"), indent=1)
print("traffic", traffic)
PY
cp profiles/ncu_traffic_C2.json $OUT/${TAG}_ncu_traffic_C2.json
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > $OUT/${TAG}_tests.log; cat $OUT/${TAG}_tests.log
python bench.py --steps 20 > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>&1
python bench.py --steps 10 --mesh-normals --no-cpu-baseline > $OUT/${TAG}_bench_mesh_normals.json 2>/dev/null
python bench.py --steps 3 --no-cpu-baseline --timeline $OUT/${TAG}_timeline_graph_replay.txt > /dev/null 2>&1
python bench.py --config C5 --steps 3 > $OUT/${TAG}_render_c5_n1.json 2> /dev/null
python bench.py --config C3 --steps 3 > $OUT/${TAG}_render_c3_n1.json 2> /dev/null
python bench.py --impl reference --config C5 --steps 1 --warmup 0 > $OUT/${TAG}_render_c5_reference_arm.json 2>&1
timeout 120 python tools/bench_bvh.py > $OUT/${TAG}_bvh.json
python - $TAG <<'PY'
import json, sys
tag = sys.argv[1]
for f in ("bench_default", "bench_mesh_normals", "render_c5_n1", "render_c3_n1"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s_%s.json" % (tag, f)) if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "%.4g" % d["value"], (d.get("roofline") or {}).get("frac"), (d.get("step_roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
PY
