"""Per-SASS-instruction sample timeline of one kernel from an .ncu-rep (read here, no GPU): address, samples, share,
dominant stall reason, instruction text -- in program order, so that the latency chain of a loop can be read off.
usage: python tools/ncu_sass_timeline.py rep.ncu-rep [min_share_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
recs = []
for r in rows:
    if len(r) > 2 and ("Address" in r or "Source" in r) and "# Samples" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        recs.append(r)
if not hdr:
    print("no source page")
    sys.exit(1)
ci = {h: i for i, h in enumerate(hdr)}
samp = ci["# Samples"]
src = ci.get("Source", 1)
addr = ci.get("Address", 0)
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
inst_col = ci.get("Instructions Executed")
tot = sum(int(r[samp] or 0) for r in recs) or 1
print(f"total samples {tot}; columns: addr  samples  share%  cum%  inst_exec  top-stall  sass")
cum = 0
for r in recs:
    s = int(r[samp] or 0)
    cum += s
    if 100.0 * s / tot < min_share:
        continue
    best = max(stall_cols, key=lambda hc: int(r[hc[1]] or 0)) if stall_cols else ("", 0)
    bs = int(r[best[1]] or 0) if stall_cols else 0
    print(f"{r[addr][-5:]:>6s} {s:7d} {100.0 * s / tot:6.2f} {100.0 * cum / tot:6.1f} {r[inst_col] if inst_col is not None else '':>10s}  "
          f"{best[0].replace('stall_', ''):>14s}:{bs:<6d} {r[src][:90]}")
