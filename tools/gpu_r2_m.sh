#!/bin/bash
# GPU call M: packed FFMA2 in pdsc_layer_kernel / window_attention_kernel: PointDSC + pipeline parity tests, bench
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_pointdsc_gpu.py tests/test_pipeline_gpu.py tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02m_pytest.log
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02m_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02m_bench.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
except Exception as e:
    print("unreadable", e)
PY
