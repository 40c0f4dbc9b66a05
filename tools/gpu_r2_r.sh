#!/bin/bash
# round 2, GPU call R: fp8 cross-term GEMM (precision 2): GEMM tests, network parity at both precisions, bench at precision 2 and 3,
# tensor-memory read-out micro-benchmark (fixed)
mkdir -p gpurun_out
timeout 120 tools/bin/tmem_ld_bw > gpurun_out/r02r_tmem_ld_bw.json 2> gpurun_out/r02r_tmem_ld_bw.err; echo "tmem_ld_bw exit $?"
cat gpurun_out/r02r_tmem_ld_bw.json
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu -s > gpurun_out/r02r_pytest_gemm.log 2>&1; echo "gemm tests exit $?"
grep -E "relative rms|passed|failed|Error|error" gpurun_out/r02r_pytest_gemm.log | tail -12
timeout 900 python -m pytest tests/test_backbone_gpu.py -x -q -m gpu -s > gpurun_out/r02r_pytest_backbone.log 2>&1; echo "backbone tests exit $?"
grep -E "max abs err|precision|passed|failed|Error|error" gpurun_out/r02r_pytest_backbone.log | tail -14
timeout 600 python bench.py --no-matcher > gpurun_out/r02r_bench_p2.json 2> gpurun_out/r02r_bench_p2.err; echo "bench p2 exit $?"
timeout 600 python bench.py --no-matcher --precision 3 > gpurun_out/r02r_bench_p3.json 2> gpurun_out/r02r_bench_p3.err; echo "bench p3 exit $?"
python - <<'PY'
import json
for n in ("p2", "p3"):
    try:
        d = json.loads(open(f"gpurun_out/r02r_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d.get("kernels_ms_per_step"), d.get("network_gemm"))
    except Exception as e:
        print(n, "unreadable", e)
        print(open(f"gpurun_out/r02r_bench_{n}.err").read()[-1500:])
PY
