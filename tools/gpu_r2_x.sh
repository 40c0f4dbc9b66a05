#!/bin/bash
# round 2, GPU call X: the whole GPU suite, smoke, and the default bench line (both arms) on the tree with GEMM precision 2 as default
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r02x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02x_smoke.log
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02x_bench_reference.json 2> gpurun_out/r02x_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
for n in ("r02x_bench", "r02x_bench_reference"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "clocks", "e2e", "cpu_baseline", "gpu_launches", "kernels_ms_per_step")})
        if "roofline" in l: print({k: l["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic")}, l.get("roofline_config5", {}) and {k: l["roofline_config5"].get(k) for k in ("achieved", "frac")})
    except Exception as e:
        print(n, "unreadable", e)
PY
