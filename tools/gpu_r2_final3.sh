#!/bin/bash
# round 2, end-of-round evidence on the final tree (one GPU): whole GPU suite, smoke, both bench arms, then (if this box runs the
# library under ncu) the ncu traffic capture whose fingerprint bench.py checks
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_final_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 \
  && { timeout 850 python tools/ncu_traffic.py > gpurun_out/r02_ncu_traffic.out 2>&1; echo "ncu_traffic exit $?"; grep fingerprint gpurun_out/r02_ncu_traffic.out; cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json; } \
  || echo "this box crashes the library under ncu: no traffic capture"
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
for n in ("r02_final_bench", "r02_final_bench_reference"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "clocks", "e2e", "cpu_baseline", "gpu_launches", "kernels_ms_per_step")})
        if "roofline" in l: print({k: l["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic", "frac_of_burst_peak")}, l.get("roofline_config5") and {k: l["roofline_config5"].get(k) for k in ("achieved", "frac")})
    except Exception as e:
        print(n, "unreadable", e)
PY
