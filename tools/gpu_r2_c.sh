#!/bin/bash
# Round 2, third GPU call: CTA-pair GEMM (gemm_tc2_kernel) parity and A/B, pair matcher test, per-shape GEMM rates.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_match_gpu.py -m gpu -x -q > gpurun_out/r02c_pytest_gemm_match.log 2>&1; echo "gemm+match tests exit $?"; tail -3 gpurun_out/r02c_pytest_gemm_match.log
timeout 300 python tools/bench_gemm_all.py > gpurun_out/r02c_gemm_shapes_pair.txt 2>&1; echo "gemm shapes pair exit $?"
ORYON_GEMM_1CTA=1 timeout 300 python tools/bench_gemm_all.py > gpurun_out/r02c_gemm_shapes_1cta.txt 2>&1; echo "gemm shapes 1cta exit $?"
timeout 700 python -m pytest tests/test_backbone_gpu.py tests/test_pipeline_gpu.py tests/test_checkpoint_gpu.py -m gpu -x -q > gpurun_out/r02c_pytest_net.log 2>&1; echo "network tests exit $?"; tail -3 gpurun_out/r02c_pytest_net.log
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02c_bench_pair.json 2> gpurun_out/r02c_bench_pair.err; echo "bench pair exit $?"; tail -2 gpurun_out/r02c_bench_pair.err
ORYON_GEMM_1CTA=1 timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02c_bench_1cta.json 2> gpurun_out/r02c_bench_1cta.err; echo "bench 1cta exit $?"
python - <<'PY'
import json
for n in ("r02c_bench_pair", "r02c_bench_1cta"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"], l["network_gemm"])
    except Exception as e:
        print(n, "unreadable", e)
for n in ("pair", "1cta"):
    print(n)
    for line in open(f"gpurun_out/r02c_gemm_shapes_{n}.txt"):
        if line.startswith(("clip_", "_total", "swin2_fc1", "f_clipconv")):
            print("  ", line.strip()[:200])
PY
