#!/bin/bash
# GPU call H: multi-row LayerNorm for narrow rows (bit-identity check vs ORYON_LN_V1, timing), 32 pairs per network pass
mkdir -p gpurun_out
ORYON_LN_V1=1 timeout 200 python tools/attn_check.py save > gpurun_out/r02h_ln_save.json 2> gpurun_out/r02h_ln_save.err; echo "save (LN v1) exit $?"
timeout 200 python tools/attn_check.py compare > gpurun_out/r02h_ln_compare.json 2> gpurun_out/r02h_ln_compare.err; echo "compare (LN rows kernel) exit $?"; cat gpurun_out/r02h_ln_compare.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02h_bench.err
timeout 300 python bench.py --no-matcher --no-cpu-baseline --pairs-per-pass 32 > gpurun_out/r02h_bench_ppp32.json 2> gpurun_out/r02h_bench_ppp32.err; echo "bench ppp32 exit $?"; tail -2 gpurun_out/r02h_bench_ppp32.err
python - <<'PY'
import json
for n in ("r02h_bench", "r02h_bench_ppp32"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step", "gpu_launches_per_step")}, l["e2e"]["value"])
    except Exception as e:
        print(n, "unreadable", e)
PY
