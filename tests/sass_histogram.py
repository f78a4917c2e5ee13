"""Probe (not a test): SASS opcode histogram of the shipped library (cuobjdump -sass), total and, per kernel, the
tensor-core / TMEM / bulk-copy / mbarrier / cluster instructions.   usage: python tests/sass_histogram.py > profiles/sass_opcodes_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "superconductor_vae_b200", "libscvae_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
special = re.compile(r"^(UTC|LDTM|STTM|UBLK|SYNCS|UCGABAR|HMMA|LDGSTS|UTMA|MAPA|CCTL)")
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        total[op.split(".")[0]] += 1
        if special.match(op):
            cur[op] += 1
print("# cuobjdump -sass superconductor_vae_b200/libscvae_b200.so (sm_100a): opcode histogram of the end-of-round build (tests/sass_histogram.py)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit (.MULTICAST = to both CTAs of a cluster),")
print("# UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP = cp.async.bulk global -> shared (.MULTICAST = into both CTAs), UBLKPF = cp.async.bulk.prefetch.L2,")
print("# SYNCS = mbarrier ops, UCGABAR = barrier.cluster, HMMA = mma.sync (grid-barrier small-batch kernel), LDGSTS = cp.async.")
print("# No UTMALDG: operands are pre-tiled, so 1-D bulk copies suffice.")
for op, n in total.most_common(70):
    print(f"{n:8d} {op}")
tot_special = collections.Counter()
for c in per.values():
    for op, n in c.items():
        tot_special[op.split(".")[0] + (".2CTA" if ".2CTA" in op else "")] += n
print("\n# totals of the special instructions: " + ", ".join(f"{op} x{n}" for op, n in tot_special.most_common()))
print("\n# per kernel: tensor-core / TMEM / bulk-copy / mbarrier / cluster instructions")
for name, c in per.items():
    if c:
        print(f"{name}: " + ", ".join(f"{op} x{n}" for op, n in c.most_common()))
