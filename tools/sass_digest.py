"""SASS opcode digest of libicpcuda.so per kernel (cuobjdump -sass): instruction counts by class, with the opcodes that
identify the tensor / copy paths (DMMA = mma.sync f64, UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, LDGSTS = cp.async,
UBLKCP / UTMALDG = TMA).     python tools/sass_digest.py > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "icp-proposal_b200", "libicpcuda.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern = None
stats = collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        stats[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        stats[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(stats), capture_output=True, text=True).stdout.splitlines()
KEY = ["DMMA", "UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "LDGSTS", "UBLKCP", "UTMALDG", "DFMA", "DADD", "DMUL",
       "FFMA", "FFMA2", "FADD2", "FMUL2", "HMMA", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "SYNCS", "WARPSYNC", "MUFU"]
print("kernel | total | " + " ".join(KEY))
tot = collections.Counter()
for (k, c), name in zip(stats.items(), demangle):
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    print(f"{short[:70]:70s} | {sum(c.values()):6d} | " + " ".join(f"{kk}={c[kk]}" for kk in KEY if c[kk]))
    tot.update(c)
print("\nlibrary totals: " + " ".join(f"{kk}={tot[kk]}" for kk in KEY if tot[kk]))
