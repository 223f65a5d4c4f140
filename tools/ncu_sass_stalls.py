"""Stall-sample summary of one kernel from `ncu -i rep --page source --csv` (SASS view): shares of the stall reasons, samples and
executed instructions by opcode, and the hottest instruction windows.  usage: python tools/ncu_sass_stalls.py file_source.csv [nwin]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = rows[1]
col = {c: i for i, c in enumerate(hdr)}
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
def num(x):
    try: return float(x)
    except Exception: return 0.0
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
    op = m.group(1) if m else "?"
    ins.append((op, r[col["Source"]].strip(), num(r[col["# Samples"]]), num(r[col["Instructions Executed"]]), {s: num(r[col[s]]) for s in stalls}))
ts = sum(i[2] for i in ins) or 1; ti = sum(i[3] for i in ins) or 1
print(f"# {rows[0][1]}\n# samples {ts:.0f}, warp-instructions executed {ti:.4g}, SASS lines {len(ins)}")
tot = collections.Counter()
for i in ins:
    for s, v in i[4].items(): tot[s] += v
print("# stall reasons:", ", ".join(f"{s[6:]} {100*v/ts:.1f}%" for s, v in tot.most_common(9)))
byop = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for op, _, sm, ex, st in ins:
    key = "FP64" if op in ("DFMA", "DADD", "DMUL", "DSETP") else op
    byop[key][0] += sm; byop[key][1] += ex
    for s, v in st.items(): byop[key][2][s] += v
print("# by opcode: samples%, executed%, main reasons")
for op, (sm, ex, st) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"  {op:10s} {100*sm/ts:5.1f}%  {100*ex/ti:5.1f}%  " + " ".join(f"{s[6:]}={100*v/max(sm,1):.0f}%" for s, v in st.most_common(3)))
# hottest windows of 40 consecutive instructions
W = 40
best = []
for a in range(0, len(ins), W):
    sm = sum(i[2] for i in ins[a:a + W])
    best.append((sm, a))
print(f"# hottest windows of {W} SASS instructions: samples%, first line, opcode mix")
for sm, a in sorted(best, reverse=True)[:nwin]:
    mix = collections.Counter(i[0] for i in ins[a:a + W])
    print(f"  {100*sm/ts:5.1f}%  @{a:6d}  " + " ".join(f"{o}:{c}" for o, c in mix.most_common(6)))
