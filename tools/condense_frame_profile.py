"""ncu --csv launch list of tools/profile_frame.py -> (a) profiles-style launch list of the LAST frame with per-kernel shares,
(b) the kernel,dram_read_MB,dram_write_MB,time_us table bench.py reads (roofline.traffic).
usage: python tools/condense_frame_profile.py launches.csv out_list.txt out_traffic.csv"""
import csv, re, sys
from collections import OrderedDict
src, out_list, out_csv = sys.argv[1:4]
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = OrderedDict()
for r in csv.DictReader(lines[start:]):
    i = int(r["ID"])
    d = rows.setdefault(i, {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""), "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    sc = {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    d[r["Metric Name"]] = v * sc
ids = list(rows)
first = [i for i in ids if rows[i]["name"].startswith("first_conv")]
sel = [i for i in ids if i >= first[-1]]
tot = sum(rows[i].get("gpu__time_duration.sum", 0.0) for i in sel)
per = OrderedDict()
with open(out_list, "w") as f:
    for k, i in enumerate(sel):
        d = rows[i]
        t = d.get("gpu__time_duration.sum", 0.0)
        f.write(f"{k:3d} {d['name'][:46]:46s} {d['grid']:16s} {t:9.1f} us  dram {(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:7.0f} MB"
                f"  L2->SM {d.get('lts__t_sectors_srcunit_tex_op_read.sum', 0) * 32 / 1e6:7.0f} MB  tensor pipe {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f} %\n")
        key = re.sub(r"<.*", "", d["name"])
        per[key] = per.get(key, 0.0) + t
    f.write("\n")
    for k, t in sorted(per.items(), key=lambda kv: -kv[1]):
        f.write(f"{k:44s} {t / 1e3:7.3f} ms {100 * t / tot:5.1f}%\n")
    f.write(f"total {tot / 1e3:.3f} ms (serialised, cold-cache launches under ncu)\n")
with open(out_csv, "w") as f:
    f.write("kernel,dram_read_MB,dram_write_MB,time_us\n")
    for i in sel:
        d = rows[i]
        if d["name"].startswith(("conv_", "first_conv")):
            f.write(f"{d['name']},{d.get('dram__bytes_read.sum', 0) / 1e6:.1f},{d.get('dram__bytes_write.sum', 0) / 1e6:.1f},{d.get('gpu__time_duration.sum', 0):.1f}\n")
