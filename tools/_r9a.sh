(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r9a_tests.log; cat gpurun_out/r9a_tests.log
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r9a_bench.json 2> gpurun_out/r9a_bench.err
MIRRES_VIS_TAGS=0 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r9a_bench_notags.json 2>/dev/null
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r9a_bench2.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench","bench_notags","bench2"):
    try:
        d=json.loads([l for l in open("gpurun_out/r9a_%s.json"%f) if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["roofline"]["frac"], d["step_roofline"]["frac"], d["kernel_ms_per_step"])
    except Exception as e: print(f,"FAILED",e)
PY
