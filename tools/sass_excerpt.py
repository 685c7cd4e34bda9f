"""Counts the tensor-core / TMA / TMEM instructions per kernel in the SASS of libvcg_b200.so.
usage: cuobjdump -sass vae-cyclegan-implementation_b200/libvcg_b200.so | python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

OPS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HGMMA", "HMMA", "REDG", "SYNCS")
cur, counts, first = None, collections.OrderedDict(), {}
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur], first[cur] = collections.Counter(), {}
        continue
    if cur is None:
        continue
    for op in OPS:
        if re.search(r"\b" + op, line):
            counts[cur][op] += 1
            first[cur].setdefault(op, re.sub(r"\s+", " ", line.strip())[:140])
print("# cuobjdump -sass libvcg_b200.so (sm_100a): tensor-core / TMA / TMEM instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma (bf16), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM -> registers),")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDG = red.global.add (split-K weight gradients),")
print("# HMMA = legacy mma.sync (only wgrad_thin_mma_kernel, the small-map fallback of the 64->3 layer); no HGMMA (sm_90a wgmma)")
for k, c in counts.items():
    if not any(c[o] for o in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA", "HGMMA", "UTMASTG")):
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0][-64:]
    print(f"{name:66s} " + "  ".join(f"{op}={n}" for op, n in c.items()))
print()
for want in ("conv_tc2_kernel", "wgrad_tc2_kernel"):
    for k in counts:
        if want in k:
            print(f"# first occurrence of each of them in {want}:")
            for op, l in first[k].items():
                print("    ", l)
