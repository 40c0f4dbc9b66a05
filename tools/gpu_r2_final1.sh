#!/bin/bash
# round 2, final evidence 1 (one GPU): ncu traffic capture of the matcher kernels (-> profiles/ncu_traffic.json), ncu --set full of the
# pair GEMM at precision 2, ncu launch list of one bench step
mkdir -p gpurun_out; nvidia-smi --query-gpu=driver_version --format=csv,noheader
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu (driver above): giving up early"; exit 3; }
timeout 850 python tools/ncu_traffic.py > gpurun_out/r02_ncu_traffic.out 2>&1; echo "ncu_traffic exit $?"; tail -5 gpurun_out/r02_ncu_traffic.out
bash tools/gpu_r2_y.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_final_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-matcher --no-affinity > gpurun_out/r02_final_ncu_launches.log 2>&1; echo "launch list exit $?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_final_launches_bench.csv", errors="ignore")))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[i0 + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = r[kn].split("(")[0][:70]
    agg[name][0] += v; agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
out = [f"{v[0]/1e3:12.1f} us {v[1]:5d} x {v[0]/v[1]/1e3:9.1f} {100*v[0]/tot:5.1f}%  {k}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])]
open("gpurun_out/r02_final_launches_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[:25]))
PY
