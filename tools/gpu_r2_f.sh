#!/bin/bash
# GPU call F: attention with P in tensor memory (.ts MMA): check vs materialised path, timeline, bench
mkdir -p gpurun_out
timeout 200 python tools/attn_check.py save > gpurun_out/r02f_attn_save.json 2> gpurun_out/r02f_attn_save.err; echo "attn save exit $?"; cat gpurun_out/r02f_attn_save.json; tail -3 gpurun_out/r02f_attn_save.err
ORYON_ATTN_MATERIALIZED=1 timeout 200 python tools/attn_check.py compare > gpurun_out/r02f_attn_compare.json 2> gpurun_out/r02f_attn_compare.err; echo "attn compare exit $?"; cat gpurun_out/r02f_attn_compare.json
ORYON_ATTN_DEBUG=1 timeout 200 python tools/bench_backbone.py --pairs 16 --steps 1 > gpurun_out/r02f_attn_dbg.json 2> gpurun_out/r02f_attn_dbg.err; echo "attn dbg exit $?"; grep -A 20 "attn_tc dbg" gpurun_out/r02f_attn_dbg.err | head -30; cat gpurun_out/r02f_attn_dbg.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02f_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
except Exception as e:
    print("unreadable", e)
PY
