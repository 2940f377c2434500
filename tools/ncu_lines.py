"""Aggregate an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hk = next(k for k, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hk]; ci = {}
for i, h in enumerate(hdr): ci.setdefault(h, i)
iS, iI, iT = ci["Warp Stall Sampling (All Samples)"], ci["Instructions Executed"], ci["Thread Instructions Executed"]
agg = collections.OrderedDict(); f = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if len(r) <= iI or r[0] in ("", "Line No", "Function Name"): continue
    try: agg[(f, r[0], r[1].strip()[:105])] = [float(r[iS] or 0), float(r[iI] or 0), float(r[iT] or 0)]
    except ValueError: pass
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"samples {ts:.0f} inst {ti:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s L%-4s %5.1f%% samp %5.1f%% inst %4.1f thr | %s" % (k[0][:16], k[1], v[0] / ts * 100, v[1] / ti * 100, v[2] / max(v[1], 1), k[2]))
