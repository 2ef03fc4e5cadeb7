"""Condense an ncu --csv launch list (gpu__time_duration.sum [+ dram__bytes_*]) into per-kernel totals.
usage: python tools/summarize_launches.py launches.csv [first_id last_id] [--list]"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
args = [a for a in sys.argv[2:] if not a.startswith("--")]
lo, hi = (int(args[0]), int(args[1])) if len(args) >= 2 else (0, 1 << 60)
rows = {}
with open(path, newline="") as f:
    lines = f.read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(lines[start:]):
    i = int(r["ID"])
    if not lo <= i <= hi:
        continue
    d = rows.setdefault(i, {"name": r["Kernel Name"], "grid": r["Grid Size"], "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    m, u = r["Metric Name"], r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if m == "gpu__time_duration.sum":
        d["t"] = v * scale
    elif m == "dram__bytes_read.sum":
        d["rd"] = v * scale
    elif m == "dram__bytes_write.sum":
        d["wr"] = v * scale


def short(n):
    n = re.sub(r"\(.*", "", n)
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return n[:64]


if "--list" in sys.argv:
    for i, d in sorted(rows.items()):
        print(f"{i:5d} {short(d['name']):64s} {d['grid']:16s} {d['t']:9.1f} us {(d['rd'] + d['wr']) / 1e6:9.1f} MB")
tot = OrderedDict()
for d in rows.values():
    k = short(d["name"])
    a = tot.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["t"]
    a[2] += d["rd"] + d["wr"]
total = sum(a[1] for a in tot.values())
for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    gbs = a[2] / a[1] / 1e3 if a[1] > 0 else 0.0
    print(f"{k:64s} {a[0]:5d} {a[1]:9.1f} us {100 * a[1] / total:5.1f}%  {a[2] / 1e6:9.1f} MB  {gbs:7.1f} GB/s")
print(f"total {total:.1f} us over {len(rows)} launches, {sum(a[2] for a in tot.values()) / 1e9:.3f} GB DRAM")
