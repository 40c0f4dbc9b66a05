#!/bin/bash
# after comment-only edits of csrc/: the ncu traffic capture again (source fingerprint), the matcher / GEMM / backbone tests, one bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 850 python tools/ncu_traffic.py > gpurun_out/r02_ncu_traffic.out 2>&1; echo "ncu_traffic exit $?"; grep fingerprint gpurun_out/r02_ncu_traffic.out; cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_gemm_gpu.py tests/test_backbone_gpu.py -x -q -m gpu > gpurun_out/r02_final4_pytest.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02_final4_pytest.log
timeout 900 python bench.py > gpurun_out/r02_final4_bench.json 2> gpurun_out/r02_final4_bench.err; echo "bench exit $?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_final4_bench.json").read().strip().splitlines()[-1])
print({k: l.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")}, l["e2e"]["value"])
print({k: l["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic", "frac_of_burst_peak")})
PY
