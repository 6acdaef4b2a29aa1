"""per-layer table of a `ncu --set full` capture of the 46 conv launches of one batch (stem + maxpool and the decoder tail are
fused launches; gpurun_out/prof_convs_raw.csv from scripts/gpu_profile.sh) -> profiles/*.txt"""
import csv
import sys

LAYERS = ["stem+pool"]
for li, n in ((1, 3), (2, 4), (3, 6), (4, 3)):
    for b in range(n):
        if b == 0 and li > 1:
            LAYERS += [f"layer{li}.0.downsample", f"layer{li}.0.conv1", f"layer{li}.0.conv2"]
        else:
            LAYERS += [f"layer{li}.{b}.conv1", f"layer{li}.{b}.conv2"]
for i in range(4):
    LAYERS += [f"dec{i}.conv1", f"dec{i}.conv2"]
LAYERS += ["dec4.conv1", "tail(dec4.conv2+head)"]      # 46 launches: stem + maxpool and the decoder tail are fused launches

src, dst, title = sys.argv[1:4]
tiles = int(sys.argv[4]) if len(sys.argv) > 4 else 135
rows = list(csv.reader(open(src)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default="nan"):
    i = col.get(name)
    return r[i] if i is not None and r[i] not in ("", "n/a") else default
out = [f"# {title}",
       "# per launch (consecutive launches of one step re-ordered to start at the stem); ncu times are cold-cache / serialised (compare",
       "# shares); traffic = dram read + write;",
       "# tensor%act = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       f"{'layer':22s} {'kernel':38s} {'us':>7s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'dram%':>6s} {'lts%':>6s} {'l1tex%':>6s} {'tensor%act':>10s} {'regs':>5s} {'grid':>6s}"]
tot_us = tot_rd = tot_wr = 0.0
data = rows[2:][:len(LAYERS)]       # one period: 46 convs + head (a 48th captured launch repeats the first)
stem_at = next(i for i, r in enumerate(data) if "conv_stem" in r[col["Kernel Name"]])
data = data[stem_at:] + data[:stem_at]        # the capture window may straddle two consecutive batches: start at the stem
for name, r in zip(LAYERS, data):
    k = g(r, "Kernel Name").replace("void ", "").replace("<unnamed>::", "")[:38]
    us = float(g(r, "gpu__time_duration.sum").replace(",", ""))
    rd = float(g(r, "dram__bytes_read.sum").replace(",", ""))
    wr = float(g(r, "dram__bytes_write.sum").replace(",", ""))
    ru = rows[1][col["dram__bytes_read.sum"]]
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[ru]
    wu = rows[1][col["dram__bytes_write.sum"]]
    wscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[wu]
    rd *= scale; wr *= wscale
    tot_us += us; tot_rd += rd; tot_wr += wr
    out.append(f"{name:22s} {k:38s} {us:7.1f} {rd:10.1f} {wr:10.1f} "
               f"{float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
               f"{float(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
               f"{float(g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
               f"{float(g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):10.1f} "
               f"{g(r, 'launch__registers_per_thread'):>5s} {g(r, 'launch__grid_size'):>6s}")
out.append(f"total {tot_us:.1f} us; dram read {tot_rd / 1e3:.3f} GB + write {tot_wr / 1e3:.3f} GB = {(tot_rd + tot_wr) / 1e3:.3f} GB per {tiles}-tile batch "
           f"({(tot_rd + tot_wr) / tiles:.1f} MB per tile)")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
