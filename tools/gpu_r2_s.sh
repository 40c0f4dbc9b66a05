#!/bin/bash
# round 2, GPU call S: what bounds match_tc at D = 128?  The bench's matcher region with (1) accumulators released unread, (2) query
# stages loaded once, (3) both -- timing experiments, results are garbage in modes 1-3 (results_ok false).
mkdir -p gpurun_out
for m in 0 1 2 3; do
  ORYON_MATCH_DEBUG_MODE=$m timeout 300 python bench.py --matcher-only --matcher-seconds 1.0 > gpurun_out/r02s_matcher_mode$m.json 2> gpurun_out/r02s_matcher_mode$m.err; echo "mode $m exit $?"
done
python - <<'PY'
import json
for m in range(4):
    try:
        d = json.loads(open(f"gpurun_out/r02s_matcher_mode{m}.json").read().strip().splitlines()[-1])
        mc = d.get("matcher_config2") or d
        print(m, {k: mc.get(k) for k in ("value", "ms_per_step", "results_ok", "kernels_ms_per_step")}, d.get("roofline"), (d.get("roofline_config5") or {}).get("frac"))
    except Exception as e:
        print(m, "unreadable", e, open(f"gpurun_out/r02s_matcher_mode{m}.err").read()[-800:])
PY
