"""Summarise ncu output into the small tracked files under profiles/.

    python tools/ncu_summary.py launches <launches.csv> [kernel-name filter regex]
    python tools/ncu_summary.py full <report.ncu-rep>
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def launches(path, pat=None):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("oryon::", "")[:70]
        if pat and not re.search(pat, r[ki]):
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[ui], v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values()) or 1.0
    print(f"| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
    for k, v in agg.items():
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {sum(v) / tot:.3f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        print(f"### {d.get('Kernel Name', '?')[:100]}\n\n| metric | unit | value |\n|---|---|---|")
        for i, n in enumerate(h):
            if n in KEYS:
                print(f"| {n} | {u[i]} | {v[i]} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
