#!/bin/bash
# End-of-round evidence on one B200: parity tests, smoke, the default bench line, the reference arm, the ncu launch list of the
# bench command and one `ncu --set full` capture each of the three matcher kernels.  Output under gpurun_out/final_*.
mkdir -p gpurun_out
T=${1:-"tests/test_match_gpu.py tests/test_lift_gpu.py tests/test_pipeline_gpu.py"}
timeout 900 python -m pytest $T -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_smoke.log; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference exit $?"
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/final_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-full-path --no-cpu-baseline > gpurun_out/final_ncu1.log 2>&1
for k in match_tc_kernel prep_dense2_kernel refine_rows4_kernel; do
  timeout 300 $NCU --set full --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/final_$k \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-full-path --no-cpu-baseline > gpurun_out/final_ncu_$k.log 2>&1
done
ls -la gpurun_out/final_* | awk '{print $5, $9}'
python - <<'PY'
import json
for n in ("final_bench", "final_bench_reference"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "results_ok", "clocks", "e2e", "cpu_baseline", "kernels_ms_per_step")})
        if "roofline" in l: print({k: l["roofline"][k] for k in ("achieved", "peak", "frac")})
        fp = l.get("full_path")
        if fp: print({k: fp.get(k) for k in ("pairs_per_s", "ms_per_step", "network_ms", "status", "error")})
    except Exception as e:
        print(n, "unreadable", e)
PY
