#!/bin/bash
# GPU call O: PointDSC layer kernel with deeper unrolling: parity + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointdsc_gpu.py -m gpu -x -q > gpurun_out/r02o_pytest_pdsc.log 2>&1; echo "pdsc tests exit $?"; tail -3 gpurun_out/r02o_pytest_pdsc.log
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02o_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02o_bench.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"])
except Exception as e:
    print("unreadable", e)
PY
