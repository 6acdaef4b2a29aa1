"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and a --set full report into the
small text summaries committed under profiles/."""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor"]


def launches(path: Path, out: Path, title: str):
    lines = [l for l in path.read_text().splitlines(True) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r["Metric Unit"]]
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:80]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with out.open("w") as fh:
        fh.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        fh.write(f"{'kernel':82s} {'n':>5s} {'total_ms':>9s} {'share%':>7s} {'avg_us':>9s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k:82s} {v[0]:5d} {v[1] / 1e3:9.3f} {100 * v[1] / tot:7.2f} {v[1] / v[0]:9.1f}\n")
        fh.write(f"total {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(rep: Path, out: Path, title: str):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(h, i) for i, h in enumerate(hdr) if h in WANT or h == "Kernel Name"]
    with out.open("w") as fh:
        fh.write(f"# {title}\n# ncu --set full --clock-control none --import-source on; per launch\n")
        for r in rows[2:]:
            fh.write("----\n")
            for h, i in idx:
                fh.write(f"  {h} [{units[i]}] = {r[i][:110]}\n")


if __name__ == "__main__":
    kind, src, dst, title = sys.argv[1:5]
    (launches if kind == "launches" else full)(Path(src), Path(dst), title)
