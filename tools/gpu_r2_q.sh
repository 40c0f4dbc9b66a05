#!/bin/bash
# round 2, GPU call Q: tensor-memory read-out micro-benchmark; PointDSC tests on both network forms; bench (default = fp32 PointDSC layers)
mkdir -p gpurun_out
timeout 120 tools/bin/tmem_ld_bw > gpurun_out/r02q_tmem_ld_bw.json 2> gpurun_out/r02q_tmem_ld_bw.err; echo "tmem_ld_bw exit $?"
cat gpurun_out/r02q_tmem_ld_bw.json
timeout 900 python -m pytest tests/test_pointdsc_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/r02q_pytest_pdsc.log 2>&1; echo "pdsc+pipeline tests exit $?"
tail -3 gpurun_out/r02q_pytest_pdsc.log
timeout 600 python bench.py --no-matcher > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench.json").read().strip().splitlines()[-1])
print("r02q_bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernels_ms_per_step"))
PY
