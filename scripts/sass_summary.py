"""Per-kernel count of the SASS opcodes that show what each kernel is made of (profiles/r2_sass_summary.txt):
tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, mbarrier = SYNCS, legacy tensor path = HMMA.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "resshift_b200" / "lib" / "librs_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
ops = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS",
       "LDSM", "MUFU", "ATOM", "RED", "MEMBAR", "BAR.SYNC", "SHFL", "LDG", "STG", "LDS", "STS"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    ins = m.group(1)
    for o in ops:
        if ins == o or ins.startswith(o + "."):
            if o == "UTCHMMA" and ins.startswith("UTCHMMA.2CTA"):
                continue
            counts[cur][o] += 1
            break
print(f"# SASS opcode counts per kernel of {lib.name} (cuobjdump -sass, sm_100a); columns with a zero everywhere are dropped")
used = [o for o in ops if any(c[o] for c in counts.values())]
print("kernel".ljust(58) + "".join(o.rjust(13) for o in used))
for k, c in counts.items():
    if not any(c.values()):
        continue
    print(k[:57].ljust(58) + "".join(str(c[o]).rjust(13) for o in used))
