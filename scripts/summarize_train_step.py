"""ncu launch list (csv, --metrics gpu__time_duration.sum) of the training bench -> per-kernel table of ONE step:
the launches between the last two loss evaluations (loss_partials_kernel)."""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
marks = [i for i, (k, _) in enumerate(rows) if "loss_partials_kernel" in k]
if len(marks) < 2:
    sys.exit(f"need two loss evaluations in the capture, found {len(marks)} in {len(rows)} launches")
step = rows[marks[-2]:marks[-1]]
agg = OrderedDict()
for k, us in step:
    k = k.split("(")[0][:76]
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + us)
total = sum(t for _, t in agg.values())
print(f"# one training step = {len(step)} launches, {total:.1f} us summed (serialised, cold-cache: compare SHARES)")
print(f"{'kernel':78s} {'n':>4s} {'total_us':>10s} {'share%':>7s} {'avg_us':>9s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:78s} {n:4d} {t:10.1f} {100 * t / total:7.2f} {t / n:9.1f}")
