#!/bin/bash
# per-GEMM timing inside a network pass (ORYON_GEMM_LOG=1, synchronous) + GEMM tests + backbone parity
mkdir -p gpurun_out
ORYON_GEMM_LOG=1 timeout 600 python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_gemm_log2.json 2> gpurun_out/r02_gemm_log2.err; echo "gemm log exit $?"
python - <<'PY'
import re, collections
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for line in open("gpurun_out/r02_gemm_log2.err"):
    m = re.match(r"gemm(\S*) M=(\d+) N=(\d+) K=(\d+) batch=(\d+)x(\d+).*npass=(\d+).*?([\d.]+) us\s+([\d.]+) TFLOP", line)
    if not m: continue
    key = (m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)) * int(m.group(6)), int(m.group(7)))
    agg[key][0] += float(m.group(8)); agg[key][1] += 1; agg[key][2] += float(m.group(9))
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
tot = sum(v[0] for _, v in rows)
out = [f"{v[0]/1e3:9.2f} ms {v[1]:4d} x {v[0]/v[1]:9.1f} us  {v[2]/v[1]:7.0f} TF alg  {100*v[0]/tot:5.1f}%  kind={k[0] or 'tma'} M={k[1]} N={k[2]} K={k[3]} batch={k[4]} npass={k[5]}" for k, v in rows]
open("gpurun_out/r02_gemm_log2_summary.txt", "w").write("\n".join(out) + f"\ntotal {tot/1e3:.2f} ms (all launches of the logged run: warm-up passes included)\n")
print("\n".join(out[:8])); print("total", tot / 1e3)
PY
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_backbone_gpu.py -x -q -m gpu > gpurun_out/r02_g2_pytest.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02_g2_pytest.log
timeout 600 python bench.py --no-matcher > gpurun_out/r02_g2_bench.json 2> gpurun_out/r02_g2_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_g2_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["status"], d["clocks"], d.get("kernels_ms_per_step"))
PY
