#!/bin/bash
# Round 2, first GPU call: the unrun legs of round 1 (N4 on the GPU, corrs_device='cuda', dataset mode, eager baseline), the new
# full-path bench line + reference arm, the matcher at the reference's shape on smooth maps, the ncu traffic capture.
#   gpurun --timeout 1700 -- 'bash tools/gpu_r2_a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc > gpurun_out/r02a_nproc.txt; lscpu | head -30 >> gpurun_out/r02a_nproc.txt; nvidia-smi topo -m >> gpurun_out/r02a_nproc.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02a_pytest_gpu.log; tail -4 gpurun_out/r02a_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r02a_smoke.log
timeout 400 python bench.py --eager-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02a_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err; echo "reference exit $?"
timeout 300 python tools/bench_match_refshape.py > gpurun_out/r02a_match_refshape.json 2> gpurun_out/r02a_match_refshape.err; echo "refshape exit $?"; tail -3 gpurun_out/r02a_match_refshape.err
timeout 200 python tools/gpu_dataset_mode.py > gpurun_out/r02a_dataset_mode.json 2> gpurun_out/r02a_dataset_mode.err; echo "dataset mode exit $?"; tail -c 400 gpurun_out/r02a_dataset_mode.json
timeout 600 python tools/ncu_traffic.py > gpurun_out/r02a_ncu_traffic.out 2>&1; echo "ncu traffic exit $?"; tail -5 gpurun_out/r02a_ncu_traffic.out
python - <<'PY'
import json
for n in ("r02a_bench", "r02a_bench_reference"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "e2e", "cpu_baseline", "gpu_eager_baseline", "kernels_ms_per_step", "h2d_probe", "affinity")})
        for r in ("roofline", "roofline_config5"):
            if l.get(r):
                print(r, {k: l[r][k] for k in ("achieved", "peak", "frac", "frac_of_burst_peak", "launch_ms", "kernel_share_of_step")})
        for m in ("matcher_config2", "matcher_config5"):
            if l.get(m):
                print(m, {k: l[m].get(k) for k in ("value", "ms_per_step", "region_s", "e2e", "kernels_ms_per_step", "clocks")})
    except Exception as e:
        print(n, "unreadable", e)
PY
