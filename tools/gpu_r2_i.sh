#!/bin/bash
# GPU call I: online-softmax attention: check vs materialised path (default tau and tau = 0: renewal at every increase), timeline, bench
mkdir -p gpurun_out
ORYON_ATTN_MATERIALIZED=1 timeout 200 python tools/attn_check.py save > gpurun_out/r02i_save.json 2> gpurun_out/r02i_save.err; echo "save (materialised) exit $?"
timeout 200 python tools/attn_check.py compare > gpurun_out/r02i_cmp_online.json 2> gpurun_out/r02i_cmp_online.err; echo "online tau=8 exit $?"; cat gpurun_out/r02i_cmp_online.json; tail -2 gpurun_out/r02i_cmp_online.err
ORYON_ATTN_TAU=0 timeout 200 python tools/attn_check.py compare > gpurun_out/r02i_cmp_online_tau0.json 2> gpurun_out/r02i_cmp_online_tau0.err; echo "online tau=0 exit $?"; cat gpurun_out/r02i_cmp_online_tau0.json
ORYON_ATTN_TWOPASS=1 timeout 200 python tools/attn_check.py compare > gpurun_out/r02i_cmp_twopass.json 2> gpurun_out/r02i_cmp_twopass.err; echo "two-pass exit $?"; cat gpurun_out/r02i_cmp_twopass.json
ORYON_ATTN_DEBUG=1 timeout 200 python tools/bench_backbone.py --pairs 16 --steps 1 > gpurun_out/r02i_attn_dbg.json 2> gpurun_out/r02i_attn_dbg.err; echo "attn dbg exit $?"; grep -A 20 "attn_tc dbg" gpurun_out/r02i_attn_dbg.err | head -24; cat gpurun_out/r02i_attn_dbg.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02i_bench.err
ORYON_ATTN_TWOPASS=1 timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02i_bench_twopass.json 2> gpurun_out/r02i_bench_twopass.err; echo "bench two-pass exit $?"
python - <<'PY'
import json
for n in ("r02i_bench", "r02i_bench_twopass"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
    except Exception as e:
        print(n, "unreadable", e)
PY
