#!/bin/bash
# One GPU call: matcher parity with the current defaults, then A/B bench lines of the refine kernel variants
# (ORYON_REFINE_OCC3, ORYON_REFINE_V1), the work decomposition (ORYON_MATCH_PLAN) and the first prep kernel (ORYON_PREP_V1).
mkdir -p gpurun_out
B="python bench.py --steps 20 --no-full-path --no-cpu-baseline"
timeout 300 python -m pytest tests/test_match_gpu.py -x -q -m gpu > gpurun_out/exp_pytest.log 2>&1; echo "pytest default rc=$?" | tee -a gpurun_out/exp_pytest.log
tail -3 gpurun_out/exp_pytest.log
timeout 120 $B --no-e2e > gpurun_out/exp_bench_default.json 2> gpurun_out/exp_bench_default.err
ORYON_REFINE_OCC3=1 timeout 120 $B --no-e2e > gpurun_out/exp_bench_occ3.json 2>/dev/null
ORYON_REFINE_V1=1 timeout 120 $B --no-e2e > gpurun_out/exp_bench_refv1.json 2>/dev/null
timeout 120 $B --no-e2e > gpurun_out/exp_bench_default2.json 2>/dev/null
timeout 120 $B --no-e2e --D 256 --H 240 --W 320 --B 8 --steps 10 > gpurun_out/exp_bench_c5.json 2>/dev/null
python - <<'PY'
import json
for n in ("default", "occ3", "refv1", "default2", "c5"):
    try:
        l = json.loads(open(f"gpurun_out/exp_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(l["value"], 1), round(l["ms_per_step"], 4), "ok" if l["results_ok"] else "BAD", {k: round(v, 4) for k, v in l["kernels_ms_per_step"].items()},
              round(l["roofline"]["frac"], 4), l["clocks"]["sm_mhz"], l["clocks"]["reasons"], (l.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "unreadable", e)
PY
