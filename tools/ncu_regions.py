"""Instruction share per region of inflate.cu from an ncu report: python tools/ncu_regions.py rep.ncu-rep tokens
Regions are found by the '// ---- X:' / '// D1:' style markers in the source, so the table follows the code."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; tokens = float(sys.argv[2]) if len(sys.argv) > 2 else 330e6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hk = next(k for k, r in enumerate(rows) if "Instructions Executed" in r)
ci = {}
for i, h in enumerate(rows[hk]): ci.setdefault(h, i)
iI = ci["Instructions Executed"]
src = open("zipc_b200/csrc/inflate.cu").read().splitlines()
marks = [(0, "prologue")]
for ln, t in enumerate(src, 1):
    m = re.match(r"\s*// (---- [A-Z]\w*:|D1:|D2:|E1:|E2:|lane i decodes|checks,|---- [a-z ]+:?)", t)
    if m: marks.append((ln, t.strip()[3:60]))
def region(l):
    name = marks[0][1]
    for ln, nm in marks:
        if ln <= l: name = nm
    return name
tot = collections.Counter(); f = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if len(r) <= iI or r[0] in ("", "Line No", "Function Name"): continue
    try: v = float(r[iI] or 0); l = int(r[0])
    except ValueError: continue
    tot[region(l) if f == "inflate.cu" else "inlined:" + f] += v
ti = sum(tot.values())
for k, v in tot.most_common(): print(f"{k:60s} {v / ti * 100:5.1f}%  {v / tokens:5.1f} instr/token")
