#!/bin/bash
# Round 2, second GPU call: CTA-pair matcher (cta_group::2) parity + A/B against the single-CTA kernel, pipelined test_step, whole suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_match_gpu.py -m gpu -x -q > gpurun_out/r02b_pytest_match.log 2>&1; echo "match tests exit $?"; tail -3 gpurun_out/r02b_pytest_match.log
timeout 120 python bench.py --matcher-only > gpurun_out/r02b_matcher_pair.json 2> gpurun_out/r02b_matcher_pair.err; echo "matcher pair exit $?"
ORYON_MATCH_1CTA=1 timeout 120 python bench.py --matcher-only > gpurun_out/r02b_matcher_1cta.json 2> gpurun_out/r02b_matcher_1cta.err; echo "matcher 1cta exit $?"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_match_gpu.py > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02b_pytest_gpu.log; tail -6 gpurun_out/r02b_pytest_gpu.log
timeout 400 python bench.py --no-matcher > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02b_bench.err
timeout 300 python bench.py --no-matcher --no-pipeline --no-cpu-baseline > gpurun_out/r02b_bench_nopipe.json 2> gpurun_out/r02b_bench_nopipe.err; echo "bench nopipe exit $?"
timeout 600 python tools/ncu_traffic.py > gpurun_out/r02b_ncu_traffic.out 2>&1; echo "ncu traffic exit $?"
python - <<'PY'
import json
for n in ("r02b_matcher_pair", "r02b_matcher_1cta"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, l["value"], l["ms_per_step"], l["results_ok"], l["kernels_ms_per_step"], l["clocks"], {k: l["roofline"][k] for k in ("achieved", "frac", "frac_of_burst_peak", "launch_ms")})
    except Exception as e:
        print(n, "unreadable", e)
for n in ("r02b_bench", "r02b_bench_nopipe"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "profiled_step_ms", "kernels_ms_per_step")}, l["e2e"]["value"], l.get("cpu_baseline"))
    except Exception as e:
        print(n, "unreadable", e)
try:
    t = json.load(open("gpurun_out/ncu_traffic.json"))
    for k, v in t["kernels"].items():
        print(k, v["gpu__time_duration.sum"], v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], v["dram_bytes_read"], v["dram_bytes_write"])
except Exception as e:
    print("traffic unreadable", e)
PY
