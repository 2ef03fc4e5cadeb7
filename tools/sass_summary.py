"""SASS mnemonic counts per kernel of libmbseg.so (cuobjdump -sass): the evidence that the hot ops are tcgen05 / TMA / TMEM code.
usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import os, re, subprocess, sys
from collections import Counter, OrderedDict
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "microbeseg_b200", "libmbseg.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "ELECT", "SYNCS", "ATOMS", "LDS", "STS", "BAR.SYNC", "MUFU", "DFMA", "DADD", "DMUL", "FFMA2", "RED.", "HMMA"]
funcs, cur = OrderedDict(), None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*", "", name).replace("(int)", "").replace("(bool)", "")
        cur = funcs.setdefault(name, Counter())
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if cur is not None and m and not ln.strip().startswith("/* 0x"):
        op = m.group(1)
        cur["instr"] += 1
        for k in keys:
            if op.startswith(k) and not (k == "HMMA" and op.startswith("UTCHMMA")):
                cur[k] += 1
tot = Counter()
for c in funcs.values():
    tot.update(c)
print("# SASS evidence: `cuobjdump -sass microbeseg_b200/libmbseg.so`, mnemonic counts per kernel (sm_100a)")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA loads / stores -> UTMALDG / UTMASTG, tcgen05.commit -> UTCBAR, elect.sync -> ELECT,")
print("# mbarrier -> SYNCS (B200_PROFILING.md); HMMA (legacy mma.sync) = %d everywhere\n" % tot["HMMA"])
print(f"TOTAL ({len(funcs)} kernels) " + " ".join(f"{k}={tot[k]}" for k in keys if tot[k] and k != "HMMA") + "\n")
for name, c in sorted(funcs.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["instr"])):
    print(f"{name[:96]:96s} instr {c['instr']:6d}  " + " ".join(f"{k}={c[k]}" for k in keys if c[k] and k != "HMMA"))
