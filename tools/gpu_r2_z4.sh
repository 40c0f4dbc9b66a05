#!/bin/bash
# round 2, GPU call Z4: ping-pong attention with fp8 cross terms in Q K^T: network parity tests, bench (default vs ORYON_ATTN_QK16=1 vs lockstep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backbone_gpu.py -x -q -m gpu -s > gpurun_out/r02z4_pytest_backbone.log 2>&1; echo "backbone tests exit $?"
grep -E "max abs err|ping-pong|precision|passed|failed|Error|error|timed out" gpurun_out/r02z4_pytest_backbone.log | tail -14
timeout 600 python bench.py --no-matcher > gpurun_out/r02z4_bench_pp.json 2> gpurun_out/r02z4_bench_pp.err; echo "bench pp exit $?"
ORYON_ATTN_QK16=1 timeout 600 python bench.py --no-matcher > gpurun_out/r02z4_bench_qk16.json 2> gpurun_out/r02z4_bench_qk16.err; echo "bench qk16 exit $?"
ORYON_ATTN_LOCKSTEP=1 timeout 600 python bench.py --no-matcher > gpurun_out/r02z4_bench_lockstep.json 2> gpurun_out/r02z4_bench_lockstep.err; echo "bench lockstep exit $?"
python - <<'PY'
import json
for n in ("pp", "qk16", "lockstep"):
    try:
        d = json.loads(open(f"gpurun_out/r02z4_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["status"], d["clocks"], d.get("kernels_ms_per_step"))
    except Exception as e:
        print(n, "unreadable", e)
        print(open(f"gpurun_out/r02z4_bench_{n}.err").read()[-1500:])
PY
