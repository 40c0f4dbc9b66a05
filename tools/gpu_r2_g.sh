#!/bin/bash
# GPU call G: whole GPU suite on the current tree, per-kernel launch list of a network pass, dataset-mode throughput (decode included)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02g_pytest_gpu.log; tail -4 gpurun_out/r02g_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_network_launches.csv python tools/bench_backbone.py --pairs 16 --steps 1 > gpurun_out/r02g_ncu_net.log 2>&1; echo "ncu launch list exit $?"
timeout 300 python tools/gpu_dataset_mode.py --repeat 43 --workers 8 > gpurun_out/r02g_dataset_w8.json 2> gpurun_out/r02g_dataset_w8.err; echo "dataset w8 exit $?"; cat gpurun_out/r02g_dataset_w8.json
timeout 300 python tools/gpu_dataset_mode.py --repeat 43 --workers 14 > gpurun_out/r02g_dataset_w14.json 2> gpurun_out/r02g_dataset_w14.err; echo "dataset w14 exit $?"; cat gpurun_out/r02g_dataset_w14.json
python - <<'PY'
import csv, collections
try:
    rows = list(csv.reader(open("gpurun_out/r02g_network_launches.csv", errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]; kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0][:70]
        v = float(r[mv].replace(",", "")); u = r[mu]
        v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        print(f"{us:10.1f} us {n:5d} x {us / n:8.1f}  {100 * us / tot:5.1f}%  {k}")
except Exception as e:
    print("launch list unreadable", e)
PY
