"""Per-kernel SASS evidence of the shipped library: counts of the instructions that show how each kernel moves its
bytes (128-bit global / shared accesses, warp collectives) and the explicit absence of the TMA / tensor-core ones.
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import re, subprocess, sys
from collections import Counter, OrderedDict
lib = sys.argv[1] if len(sys.argv) > 1 else "nuts333_b200/_lib/libnutsb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
pats = OrderedDict([("LDG.E.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.E.128.CONSTANT", r"\bLDG\.E\.128\.CONSTANT"), ("STG.E.128", r"\bSTG\.E(\.\w+)*\.128"),
                    ("LDS.128", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"), ("LDS.U16", r"\bLDS\.U16"), ("REDUX", r"\bREDUX"), ("MATCH", r"\bMATCH\."),
                    ("VOTE", r"\bVOTE\."), ("SHFL", r"\bSHFL\."), ("ATOMS", r"\bATOMS"), ("ATOMG/RED", r"\b(ATOMG|RED)\."), ("PRMT", r"\bPRMT"),
                    ("SHF (funnel)", r"\bSHF\."), ("CCTL/prefetch", r"\bCCTL"),
                    ("UBLKCP/UTMALDG/UTMASTG (TMA)", r"\b(UBLKCP|UTMALDG|UTMASTG)"), ("UTC*MMA/LDTM/STTM (tcgen05)", r"\b(UTC\w*MMA|LDTM|STTM)"), ("HMMA", r"\bHMMA")])
cur, rows = None, OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); rows[cur] = Counter(); continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        rows[cur]["instructions"] += 1
        for k, p in pats.items():
            if re.search(p, line): rows[cur][k] += 1
def short(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", d)
print("SASS summary of", lib, "-- cubins for", ", ".join(arch))
print("(counts of static instructions per kernel; TMA and tensor-core mnemonics are absent by design: nothing on this path is a")
print(" contraction, and the bulk-copy attempt -- commit ff86d4c, 1.66-2.22 ms against 0.89 -- was measured and rejected, DESIGN.md 4.1)\n")
keys = list(pats)
for fn, c in rows.items():
    print(short(fn), "-", c["instructions"], "instructions")
    print("    " + ", ".join(f"{k} {c[k]}" for k in keys if c[k]) + ("" if any(c[k] for k in keys) else "(none of the listed)"))
tot = Counter()
for c in rows.values(): tot.update(c)
print("\nwhole library:", ", ".join(f"{k} {tot[k]}" for k in keys))
