#!/bin/bash
# GPU call N: per-row-block accumulator hand-off in match_tc_kernel: parity, config-2 / config-5 timing, ncu capture (tensor pipe %, traffic)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_match_gpu.py -m gpu -x -q > gpurun_out/r02n_pytest_match.log 2>&1; echo "match tests exit $?"; tail -3 gpurun_out/r02n_pytest_match.log
timeout 120 python bench.py --matcher-only > gpurun_out/r02n_matcher.json 2> gpurun_out/r02n_matcher.err; echo "matcher exit $?"
timeout 600 python tools/ncu_traffic.py > gpurun_out/r02n_ncu_traffic.out 2>&1; echo "ncu traffic exit $?"
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02n_matcher.json").read().strip().splitlines()[-1])
    print(l["value"], l["ms_per_step"], l["results_ok"], l["kernels_ms_per_step"], l["clocks"], {k: l["roofline"][k] for k in ("achieved", "frac", "frac_of_burst_peak", "launch_ms")})
except Exception as e:
    print("unreadable", e)
try:
    t = json.load(open("gpurun_out/ncu_traffic.json"))
    for k, v in t["kernels"].items():
        print(k, v["gpu__time_duration.sum"], v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], v["dram_bytes_read"], v["dram_bytes_write"])
except Exception as e:
    print("traffic unreadable", e)
PY
