#!/bin/bash
# GPU call J: online attention with packed fp32x2 arithmetic: check vs materialised path, bench
mkdir -p gpurun_out
ORYON_ATTN_MATERIALIZED=1 timeout 200 python tools/attn_check.py save > gpurun_out/r02j_save.json 2> gpurun_out/r02j_save.err; echo "save (materialised) exit $?"
timeout 200 python tools/attn_check.py compare > gpurun_out/r02j_cmp_online.json 2> gpurun_out/r02j_cmp_online.err; echo "online exit $?"; cat gpurun_out/r02j_cmp_online.json; tail -2 gpurun_out/r02j_cmp_online.err
ORYON_ATTN_TAU=0 timeout 200 python tools/attn_check.py compare > gpurun_out/r02j_cmp_online_tau0.json 2> gpurun_out/r02j_cmp_online_tau0.err; echo "online tau=0 exit $?"; cat gpurun_out/r02j_cmp_online_tau0.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02j_bench.err
python - <<'PY'
import json
for n in ("r02j_bench",):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
    except Exception as e:
        print(n, "unreadable", e)
PY
