#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU minutes were spent,
# then the whole GPU suite and the default bench pair.  One B200:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_sift_kp_gpu.py -m gpu -q > gpurun_out/r2_sift_kp.log 2>&1; echo "sift_kp exit $?" | tee -a gpurun_out/r2_sift_kp.log; tail -3 gpurun_out/r2_sift_kp.log
timeout 300 python tools/gpu_dataset_mode.py > gpurun_out/r2_dataset_mode.json 2> gpurun_out/r2_dataset_mode.err; echo "dataset mode exit $?"; tail -c 600 gpurun_out/r2_dataset_mode.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2_pytest_gpu.log; tail -2 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r2_smoke.log
timeout 600 python bench.py --eager-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
for n in ("r2_bench", "r2_bench_reference"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "results_ok", "clocks", "e2e", "cpu_baseline", "gpu_eager_baseline")})
        if "roofline" in l:
            print({k: l["roofline"][k] for k in ("achieved", "peak", "frac")})
    except Exception as e:
        print(n, "unreadable", e)
PY
