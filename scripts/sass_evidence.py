"""SASS evidence for libevs.so (no GPU needed): mnemonic counts of the Blackwell tensor-core / TMEM / TMA instructions over
the whole library and per kernel.   python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "evo-ssearch_b200", "libevs.so")
WATCH = ("UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "CREDUX", "REDUX",
         "ACQBULK", "SYNCS", "ELECT", "DFMA", "ATOMG", "LDC", "MEMBAR")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = [0, collections.Counter()]
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    per[cur][0] += 1
    for w in WATCH:
        if op == w or op.startswith(w + "."):
            total[w] += 1
            per[cur][1][w] += 1
            break
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence for libevs.so (cuobjdump -sass evo-ssearch_b200/libevs.so, sm_100a), round 2, final tree")
print("Mnemonic counts over the whole library (PTX names never appear in SASS: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM,")
print("TMA = UTMALDG / UBLKCP, tcgen05.commit = UTCBAR):")
for w, c in total.most_common():
    print(f"  {w:14s} {c}")
print("\nPer kernel (instructions; watched mnemonics):")
for (mangled, (n, cnt)), name in zip(per.items(), names):
    if n == 0:
        continue
    short = re.sub(r"\s+", " ", name)
    print(f"  {short[:150]}")
    print(f"      {n} instr; " + (", ".join(f"{w} {c}" for w, c in sorted(cnt.items())) or "-"))
