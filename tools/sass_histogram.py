"""SASS instruction histogram per kernel of libautobz_cuda.so (cuobjdump -sass): the static evidence for which pipes a kernel
uses - DMMA (FP64 tensor cores), DFMA/DADD/DMUL (FP64 FMA pipe), UBLKCP / UTMA* (TMA bulk copies), SHFL, BAR, MUFU, shared /
global memory accesses.  Runs without a GPU.  Usage: python tools/sass_histogram.py [library.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "autobzcore.jl_b200", "libautobz_cuda.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
COLS = ["DMMA", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "SHFL", "UBLKCP", "UTMA", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG", "LDL", "STL", "REDUX", "ATOM"]
kern = None
hist = collections.OrderedDict()
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        for c in COLS:
            if op.startswith(c):
                hist[kern][c] += 1
                break
demangled = {}
try:
    names = list(hist)
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    demangled = dict(zip(names, out))
except Exception:
    pass
print(f"# SASS histogram of {os.path.basename(so)} (sm_100a), static instruction counts per kernel; cuobjdump -sass")
print("# " + " ".join(f"{c:>6s}" for c in ["total"] + COLS) + "  kernel")
for k, h in hist.items():
    name = demangled.get(k, k)
    name = re.sub(r"\((int|bool|unsigned int)\)", "", name)           # template arguments print as (int)4
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("abz::", "")
    print("  " + " ".join(f"{h.get(c, 0):6d}" for c in ["total"] + COLS) + "  " + name)
