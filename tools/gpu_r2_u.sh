#!/bin/bash
# round 2, GPU call U: matcher with the unit-boundary commits deferred past the next unit's first MMA: parity tests + matcher regions
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_match_gpu.py -x -q -m gpu > gpurun_out/r02u_pytest_match.log 2>&1; echo "match tests exit $?"
tail -3 gpurun_out/r02u_pytest_match.log
timeout 300 python bench.py --matcher-only > gpurun_out/r02u_matcher.json 2> gpurun_out/r02u_matcher.err; echo "matcher exit $?"
timeout 300 python bench.py --matcher-only --matcher-config 5 > gpurun_out/r02u_matcher_c5.json 2> gpurun_out/r02u_matcher_c5.err; echo "matcher c5 exit $?"
python - <<'PY'
import json
for n in ("r02u_matcher", "r02u_matcher_c5"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("workload"), d["value"], d["ms_per_step"], d["results_ok"], d["clocks"], d["kernels_ms_per_step"], {k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst_peak")})
    except Exception as e:
        print(n, "unreadable", e, open(f"gpurun_out/{n}.err").read()[-600:])
PY
