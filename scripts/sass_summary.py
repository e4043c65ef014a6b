#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libgligen_b200.so (cuobjdump -sass): the evidence that the hot kernels are
Blackwell-native (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG,
legacy mma.sync -> HMMA).   python scripts/sass_summary.py > profiles/r2_sass_mnemonics.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gligen_b200", "libgligen_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS", "HMMA", "MUFU", "LDGSTS", "SYNCS"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void glg::", "").replace("glg::", "")
        per[cur] = collections.Counter()
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
print("# SASS mnemonics per kernel of libgligen_b200.so (sm_100a, `cuobjdump -sass`)\n")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st,")
print("HMMA = legacy mma.sync (only the fallback attention kernel `attn_fwd_kernel`, off the hot path since round 2).\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA", "UTMASTG")) and c["_total"] < 400:
        continue
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    tot.update(c)
print("| **all listed** | %d | " % tot["_total"] + " | ".join(str(tot[k]) for k in KEYS) + " |")
