#!/usr/bin/env python
"""Per-kernel counts of the Blackwell tensor-core / TMEM / TMA SASS mnemonics in libmpinets_b200.so (cuobjdump -sass), the static
evidence that the kernels are tcgen05 / TMA kernels:  python scripts/sass_counts.py > profiles/r2_sass_counts.txt
  UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc,
  UTMALDG = cp.async.bulk.tensor (TMA load), SYNCS = mbarrier ops, REDUX = redux.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mpinets_b200", "libmpinets_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "SYNCS", "REDUX", "HMMA", "FFMA")
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
print("# cuobjdump -sass mpinets_b200/libmpinets_b200.so : instruction counts per kernel (sm_100a)")
print("%-64s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in MN))
for name, c in counts.items():
    if any(c[k] for k in MN[:6]) or "--all" in sys.argv:
        print("%-64s %7d " % (name[:64], c["total"]) + " ".join("%8d" % c[k] for k in MN))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("%-64s %7d " % ("TOTAL (all %d kernels)" % len(counts), tot["total"]) + " ".join("%8d" % tot[k] for k in MN))
