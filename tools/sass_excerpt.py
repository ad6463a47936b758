#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UBLKCP / UBLKPF, tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA (must be 0).

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt        (build container: cuobjdump, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dist_b200", "libdistb200.so")
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCCP", "SYNCS", "UCGABAR", "MUFU.TANH", "MUFU.EX2",
       "REDG", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|distb200::|<unnamed>::", "", name).split("(")[0]
        counts.setdefault(kern, collections.Counter())
        continue
    if kern is None or "/*" not in line:
        continue
    ins = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", line)
    for p in PAT:
        if re.search(r"(^|\s|@!?U?P\d\s+)" + re.escape(p) + r"(\.|\s|;)", ins):
            if p == "UTCHMMA" and "UTCHMMA.2CTA" in ins:
                continue
            counts[kern][p] += 1
            total[p] += 1
            break
print("# cuobjdump -sass dist_b200/libdistb200.so (sm_100a): occurrences per kernel of the Blackwell-native mnemonics")
print("# UTCHMMA[.2CTA] = tcgen05.mma (cta_group::1 / ::2), LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA),")
print("# UBLKPF = cp.async.bulk.prefetch.L2, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR = barrier.cluster; HMMA (mma.sync) must not appear")
print("%-64s %s" % ("kernel", "  ".join("%s" % p for p in PAT)))
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print("%-64s %s" % (k[:64], "  ".join("%*d" % (len(p), c[p]) for p in PAT)))
print("%-64s %s" % ("TOTAL", "  ".join("%*d" % (len(p), total[p]) for p in PAT)))
assert total["HMMA"] == 0, "legacy mma.sync code found"
