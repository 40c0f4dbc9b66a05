#!/bin/bash
# round 2, GPU call V: A/B of the commit placement in match_tc on one box: configs 2 and 5, deferred (default) vs at the boundary (mode 4), twice each
mkdir -p gpurun_out
for rep in 1 2; do
for m in 0 4; do
  for c in 2 5; do
    ORYON_MATCH_DEBUG_MODE=$m timeout 300 python bench.py --matcher-only --matcher-config $c --matcher-seconds 1.5 > gpurun_out/r02v_c${c}_mode${m}_$rep.json 2> gpurun_out/r02v_c${c}_mode${m}_$rep.err; echo "config $c mode $m exit $?"
  done
done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_c*_mode*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["results_ok"], d["clocks"]["sm_mhz"], round(d["kernels_ms_per_step"]["match_tc"], 4), round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
