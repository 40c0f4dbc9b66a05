#!/bin/bash
# GPU call P: PointDSC network on the tcgen05 GEMM: parity on both paths, pipeline tests, bench A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pointdsc_gpu.py -m gpu -q > gpurun_out/r02p_pytest_pdsc.log 2>&1; echo "pdsc tests exit $?"; tail -15 gpurun_out/r02p_pytest_pdsc.log
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02p_bench.err
ORYON_PDSC_FP32=1 timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02p_bench_fp32.json 2> gpurun_out/r02p_bench_fp32.err; echo "bench fp32 exit $?"
python - <<'PY'
import json
for n in ("r02p_bench", "r02p_bench_fp32"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step", "gpu_launches_per_step")}, l["e2e"]["value"], l["network_gemm"]["launches_per_step"])
    except Exception as e:
        print(n, "unreadable", e)
PY
