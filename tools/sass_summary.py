"""Per-kernel SASS evidence of the Blackwell-native path: counts of tcgen05 (UTC*MMA), tensor-memory (LDTM/STTM), TMA
(UTMALDG/UTMASTG/UBLKCP), legacy tensor (HMMA) and packed-fp32 (FFMA2) instructions in every kernel of
libeqxv_b200.so.   usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "eqxvision_b200", "lib", "libeqxv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "FFMA2",
        "LDGSTS", "SYNCS", "MUFU"]
per = collections.OrderedDict()
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("eqxv::", "").replace("(anonymous namespace)::", "")
        per.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["instr"] += 1
    base = op.split(".")[0]
    if base in KEYS:
        per[cur][base] += 1
    if base == "UTCHMMA" and ".2CTA" in op:
        per[cur]["UTCHMMA.2CTA"] += 1
tot = collections.Counter()
fam = collections.OrderedDict()
for k, c in per.items():
    f = re.sub(r"<.*", "", k)
    a = fam.setdefault(f, collections.Counter())
    a.update(c)
    a["variants"] += 1
    tot.update(c)
print(f"libeqxv_b200.so: {len(per)} kernels (template instantiations), {tot['instr']} SASS instructions")
print("totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
print(f"\n{'kernel family':34s} {'inst.':>5s} " + " ".join(f"{k:>12s}" for k in KEYS))
for f, c in sorted(fam.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 1000 - kv[1]["instr"]):
    print(f"{f[:34]:34s} {c['variants']:5d} " + " ".join(f"{c[k]:12d}" for k in KEYS))
