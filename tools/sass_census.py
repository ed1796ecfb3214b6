"""Counts the Blackwell-native instructions per kernel in the built library (cuobjdump -sass) and writes
profiles/<tag>_sass_census.md:  python tools/sass_census.py r02"""
import collections
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neuman_b200", "libneuman_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDCU", "F2FP.SATFINITE", "VHMNMX", "HMMA", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        op = mm.group(1)
        rows[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                rows[cur][k] += 1
                if k == "UTCHMMA" and ".2CTA" in op:
                    rows[cur]["UTCHMMA.2CTA"] += 1
demangled = subprocess.run(["cu++filt"] + list(rows), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(rows, demangled)) if len(demangled) == len(rows) else {k: k for k in rows}
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDCU", "F2FP.SATFINITE", "VHMNMX", "MUFU", "_total"]
path = os.path.join(ROOT, "profiles", f"{tag}_sass_census.md")
with open(path, "w") as f:
    f.write(f"# SASS census of `neuman_b200/libneuman_b200.so` (`cuobjdump -sass`, sm_100a; static instruction counts per kernel)\n\n")
    f.write("`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UBLKCP` = cp.async.bulk, "
            "`UTMALDG/UTMASTG` = TMA tensor load / store, `SYNCS` = mbarrier ops, `LDCU` = uniform constant loads, "
            "`F2FP.SATFINITE` = saturating fp16 packs, `VHMNMX` = packed-half max (range flag).\n\n")
    f.write("| kernel | " + " | ".join(c.replace("_total", "all instr") for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    tot = collections.Counter()
    for k, c in rows.items():
        if not any(c[x] for x in cols[:7]) and "mlp" not in k:
            continue
        nm = names[k].rsplit("(", 1)[0]
        f.write(f"| `{nm}` | " + " | ".join(str(c[x]) for x in cols) + " |\n")
    for c in rows.values():
        tot.update(c)
    f.write(f"| **library total ({len(rows)} kernels)** | " + " | ".join(str(tot[x]) for x in cols) + " |\n")
print(open(path).read())
