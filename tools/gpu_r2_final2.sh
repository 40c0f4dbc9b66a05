#!/bin/bash
# round 2, final evidence 2 (one GPU): per-GEMM timing inside a network pass (ORYON_GEMM_LOG=1, synchronous), then the ncu captures again
# (fingerprint of the library sources is path-independent now)
mkdir -p gpurun_out; nvidia-smi --query-gpu=driver_version --format=csv,noheader
python -c "from oryon_b200 import build; print('library current on this box:', build.is_current())"
ORYON_GEMM_LOG=1 timeout 600 python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_gemm_log.json 2> gpurun_out/r02_gemm_log.err; echo "gemm log exit $?"
python - <<'PY'
import re, collections
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for line in open("gpurun_out/r02_gemm_log.err"):
    m = re.match(r"gemm(\S*) M=(\d+) N=(\d+) K=(\d+) batch=(\d+)x(\d+).*npass=(\d+).*?([\d.]+) us\s+([\d.]+) TFLOP", line)
    if not m: continue
    key = (m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)) * int(m.group(6)), int(m.group(7)))
    agg[key][0] += float(m.group(8)); agg[key][1] += 1; agg[key][2] += float(m.group(9))
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
tot = sum(v[0] for _, v in rows)
out = [f"{v[0]/1e3:9.2f} ms {v[1]:4d} x {v[0]/v[1]:9.1f} us  {v[2]/v[1]:7.0f} TF alg  {100*v[0]/tot:5.1f}%  kind={k[0] or 'tma'} M={k[1]} N={k[2]} K={k[3]} batch={k[4]} npass={k[5]}" for k, v in rows]
open("gpurun_out/r02_gemm_log_summary.txt", "w").write("\n".join(out) + f"\ntotal {tot/1e3:.2f} ms (all launches of the logged run: warm-up passes included)\n")
print("\n".join(out[:16])); print("total", tot / 1e3)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 850 python tools/ncu_traffic.py > gpurun_out/r02_ncu_traffic.out 2>&1; echo "ncu_traffic exit $?"; grep fingerprint gpurun_out/r02_ncu_traffic.out
