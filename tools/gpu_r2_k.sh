#!/bin/bash
# GPU call K: cluster-of-4 GEMM with A multicast: parity, per-shape A/B against CTA pairs, network equality, bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r02k_pytest_gemm.log 2>&1; echo "gemm tests exit $?"; tail -4 gpurun_out/r02k_pytest_gemm.log
ORYON_GEMM_LOG=1 timeout 120 python -c "
import torch, sys
sys.path.insert(0, '.')
from oryon_b200 import ops
A = torch.randn(18464, 1024, device='cuda'); W = torch.randn(3072, 1024, device='cuda') * 0.03
for _ in range(3): ops.linear(A, W)
torch.cuda.synchronize()
" 2>&1 | grep -E "quad|pair" | tail -4
timeout 300 python tools/bench_gemm_all.py > gpurun_out/r02k_gemm_shapes_quad.txt 2>&1; echo "gemm shapes quad exit $?"
ORYON_GEMM_QUAD=0 timeout 300 python tools/bench_gemm_all.py > gpurun_out/r02k_gemm_shapes_pair.txt 2>&1; echo "gemm shapes pair exit $?"
ORYON_GEMM_QUAD=0 timeout 200 python tools/attn_check.py save > gpurun_out/r02k_save.json 2> gpurun_out/r02k_save.err; echo "save (pairs) exit $?"
timeout 200 python tools/attn_check.py compare > gpurun_out/r02k_cmp.json 2> gpurun_out/r02k_cmp.err; echo "compare (quad) exit $?"; cat gpurun_out/r02k_cmp.json
timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02k_bench.err
ORYON_GEMM_QUAD=0 timeout 300 python bench.py --no-matcher --no-cpu-baseline > gpurun_out/r02k_bench_pair.json 2> gpurun_out/r02k_bench_pair.err; echo "bench pair exit $?"
python - <<'PY'
import json
for n in ("r02k_bench", "r02k_bench_pair"):
    try:
        l = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, {k: l.get(k) for k in ("value", "ms_per_step", "status", "clocks", "kernels_ms_per_step")}, l["e2e"]["value"], l["network_gemm"]["frac_tensor_pipe"])
    except Exception as e:
        print(n, "unreadable", e)
for n in ("quad", "pair"):
    print(n)
    for line in open(f"gpurun_out/r02k_gemm_shapes_{n}.txt"):
        if line.startswith(("clip_qkv", "clip_out", "clip_fc", "clip_proj", "_total", "f_clipconv", "swin2_fc1")):
            print("  ", line.strip()[:200])
PY
