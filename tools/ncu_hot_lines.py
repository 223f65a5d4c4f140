"""Per-source-line hot spots of one kernel from an ncu report: instructions executed and stall samples.
usage: python tools/ncu_hot_lines.py report.ncu-rep <kernel regex> [top]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source=sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
def num(x):
    try: return float(x)
    except Exception: return 0.0
lines = []
i = 0
fname = ""
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        col = {c: j for j, c in enumerate(r)}
        i += 1
        while i < len(rows) and rows[i] and rows[i][0] not in ("File Path", "Function Name"):
            q = rows[i]
            if len(q) > col["# Samples"]:
                lines.append((fname, q[0], q[1].strip()[:100], num(q[col["Instructions Executed"]]), num(q[col["# Samples"]]),
                              {k: num(q[col[k]]) for k in col if k.startswith("stall_") and "Not Issued" not in k and len(q) > col[k]}))
            i += 1
        continue
    i += 1
ti = sum(l[3] for l in lines) or 1; ts = sum(l[4] for l in lines) or 1
print(f"total warp-instructions {ti:.3g}, samples {ts:.0f}")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    st = sorted(l[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{l[0]}:{l[1]:>4} inst {100*l[3]/ti:5.1f}% samp {100*l[4]/ts:5.1f}%  {' '.join(f'{k[6:]}={v:.0f}' for k, v in st if v)}  | {l[2]}")
