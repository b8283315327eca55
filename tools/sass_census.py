"""SASS census of the in-tree library (no GPU needed): which Blackwell-only instructions each kernel contains.
    python tools/sass_census.py > profiles/rNN_sass_mnemonics.txt
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, UTMALDG = TMA tensor load
(cp.async.bulk.tensor), UBLKCP = 1-D bulk copy (cp.async.bulk), SYNCS = mbarrier, UCGABAR = cluster barrier, REDG = global reduction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "epnet_b200", "libepnet_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTCCP", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
         "REDUX", "CREDUX", "REDG", "ATOMG", "ATOMS", "LDGSTS")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
total = collections.Counter()
for c in per.values():
    total.update(c)
print("# SASS census of epnet_b200/libepnet_b200.so (sm_100a), tools/sass_census.py")
print("total: " + ", ".join("%s %d" % (k, total[k]) for k in WATCH if total[k]))
print()
for name, c in per.items():
    hits = ", ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
    if hits:
        print("%-70s %6d instr  %s" % (name[:70], sum(c.values()), hits))
