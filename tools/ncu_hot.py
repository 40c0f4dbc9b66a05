"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        n = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    top = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    data.append((n, r[ci["Source"]], top, r[ci["Instructions Executed"]]))
tot = sum(d[0] for d in data)
print("total samples", tot)
for n, src, top, ex in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{n:7d} {100 * n / tot:5.1f}%  exec={ex:>9}  {src[:110]:110s} {top}")
