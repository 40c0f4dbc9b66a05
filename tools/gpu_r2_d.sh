#!/bin/bash
# Round 2, GPU call D: 16-warp fused attention (check vs materialised path, timing), GEMM precision sweep.
mkdir -p gpurun_out
timeout 200 python tools/attn_check.py save > gpurun_out/r02d_attn_save.json 2> gpurun_out/r02d_attn_save.err; echo "attn save exit $?"; cat gpurun_out/r02d_attn_save.json
ORYON_ATTN_MATERIALIZED=1 timeout 200 python tools/attn_check.py compare > gpurun_out/r02d_attn_compare.json 2> gpurun_out/r02d_attn_compare.err; echo "attn compare exit $?"; cat gpurun_out/r02d_attn_compare.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02d_bench.err
timeout 300 python tools/gemm_precision_sweep.py > gpurun_out/r02d_gemm_precision_sweep.json 2> gpurun_out/r02d_gemm_precision_sweep.err; echo "sweep exit $?"; cat gpurun_out/r02d_gemm_precision_sweep.json
python - <<'PY'
import json
for n in ("r02d_bench",):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
    except Exception as e:
        print(n, "unreadable", e)
PY
