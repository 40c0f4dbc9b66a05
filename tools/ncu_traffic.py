#!/usr/bin/env python
"""Regenerates profiles/ncu_traffic.json: DRAM bytes per launch of the matcher kernels, from one `ncu --set full` capture of the
bench's own config-2 command (`bench.py --matcher-only`), stamped with the fingerprint of the sources the library was built from.
`bench.py` prints these numbers as `roofline.traffic` only while the fingerprint still matches the running library (a stale capture
reads as null).  Run on a B200, one GPU:

    gpurun --timeout 900 -- 'python tools/ncu_traffic.py'        # also leaves gpurun_out/r02_match_tc.ncu-rep + a launch list
"""
import csv
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def main():
    from oryon_b200 import build as _build
    os.makedirs(OUT, exist_ok=True)
    rep = os.path.join(OUT, "r02_match_kernels")
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-f", "-o", rep,
           "-k", "regex:match_tc2?_kernel|prep_dense2_kernel|refine_rows4_kernel", "--launch-skip", "9", "--launch-count", "3",
           sys.executable, os.path.join(ROOT, "bench.py"), "--matcher-only", "--matcher-seconds", "0.01"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    open(os.path.join(OUT, "r02_ncu_traffic.log"), "w").write(p.stdout[-4000:] + "\n" + p.stderr[-4000:])
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, timeout=300).stdout
    open(os.path.join(OUT, "r02_match_kernels_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
    units = rows[1]
    kernels = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("<")[0].split("(")[0].replace("void ", "").strip()
        name = {"match_tc2_kernel": "match_tc_kernel"}.get(name, name)      # the pair kernel is the tensor-core pass bench.py reports

        def val(metric):
            if metric not in col:
                return None
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]].lower()
            if metric.startswith("dram__bytes"):
                v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return v
        kernels[f"{name}:config2"] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                                      **{m: val(m) for m in want[2:]}, "duration_unit": units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else None}
    rec = {"fingerprint": _build._fingerprint(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "command": " ".join(cmd[:-5] + ["python", "bench.py", "--matcher-only", "--matcher-seconds", "0.01"]), "kernels": kernels}
    with open(os.path.join(OUT, "ncu_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
