#!/bin/bash
# ncu launch list of network passes (32 pairs, precision 2) -> per-kernel totals
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_network_launches.csv \
  python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_network_launches.log 2>&1; echo "launch list exit $?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_network_launches.csv", errors="ignore")))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[i0 + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    if r[mu] == "ns": v /= 1e3
    elif r[mu] == "ms": v *= 1e3
    name = r[kn].split("(")[0][:70]
    agg[name][0] += v; agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
out = [f"{v[0]:12.1f} us {v[1]:5d} x {v[0]/v[1]:9.1f} {100*v[0]/tot:5.1f}%  {k}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])]
open("gpurun_out/r02_network_launches_summary.txt", "w").write("\n".join(out) + f"\ntotal {tot/1e3:.1f} ms over all passes of the run (3 network passes of 32 pairs + weight loading)\n")
print("\n".join(out[:40]))
PY
