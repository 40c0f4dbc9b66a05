#!/bin/bash
# One GPU call: matcher parity with the current defaults, then A/B bench lines of the work-decomposition kinds
# (ORYON_MATCH_PLAN) and of the prep / refine kernel versions (ORYON_PREP_V1, ORYON_REFINE_V1).
mkdir -p gpurun_out
B="python bench.py --steps 20 --no-full-path --no-cpu-baseline"
python -m pytest tests/test_match_gpu.py -x -q -m gpu > gpurun_out/exp_pytest.log 2>&1; echo "pytest default rc=$?" | tee -a gpurun_out/exp_pytest.log
tail -3 gpurun_out/exp_pytest.log
$B > gpurun_out/exp_bench_default.json 2> gpurun_out/exp_bench_default.err
ORYON_MATCH_PLAN=whole $B --no-e2e > gpurun_out/exp_bench_whole.json 2>/dev/null
ORYON_MATCH_PLAN=contiguous $B --no-e2e > gpurun_out/exp_bench_contiguous.json 2>/dev/null
$B --no-e2e > gpurun_out/exp_bench_default2.json 2>/dev/null
ORYON_PREP_V1=1 ORYON_REFINE_V1=1 $B --no-e2e > gpurun_out/exp_bench_v1kernels.json 2>/dev/null
$B --no-e2e --D 256 --H 240 --W 320 --B 8 --steps 10 > gpurun_out/exp_bench_c5.json 2>/dev/null
ORYON_MATCH_PLAN=whole $B --no-e2e --D 256 --H 240 --W 320 --B 8 --steps 10 > gpurun_out/exp_bench_c5whole.json 2>/dev/null
python - <<'PY'
import json
for n in ("default", "whole", "contiguous", "default2", "v1kernels", "c5", "c5whole"):
    try:
        l = json.loads(open(f"gpurun_out/exp_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(l["value"], 1), "ok" if l["results_ok"] else "BAD", {k: round(v, 4) for k, v in l["kernels_ms_per_step"].items()},
              round(l["roofline"]["frac"], 4), l["clocks"]["sm_mhz"], l["clocks"]["reasons"], (l.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "unreadable", e)
PY
