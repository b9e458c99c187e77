#!/usr/bin/env python
"""Aggregate an ncu source page by (file, line range) phases.  usage: ncu_phase_table.py rep denom file:lo-hi=name ..."""
import csv, io, subprocess, sys
rep = sys.argv[1]; denom = float(sys.argv[2])
specs = []
for a in sys.argv[3:]:
    loc, name = a.split("=")
    f, rng = loc.split(":")
    lo, hi = (int(x) for x in rng.split("-"))
    specs.append((f, lo, hi, name))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; acc = {}; ti = ts = 0
other = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:]))
    try: ln = int(r[0]); n = int(d["Instructions Executed"]); s = int(d["# Samples"])
    except Exception: continue
    ti += n; ts += s
    for f, lo, hi, name in specs:
        if cur == f and lo <= ln <= hi:
            a = acc.setdefault(name, [0, 0]); a[0] += n; a[1] += s; break
    else:
        a = other.setdefault(cur, [0, 0]); a[0] += n; a[1] += s
print(f"total {ti/denom:.2f} warp-instr per unit, {ts} stall samples")
for name, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{n/denom:7.2f} {100*n/ti:5.1f}%  stalls {100*s/ts:5.1f}%  {name}")
for name, (n, s) in sorted(other.items(), key=lambda kv: -kv[1][0]):
    print(f"{n/denom:7.2f} {100*n/ti:5.1f}%  stalls {100*s/ts:5.1f}%  [file] {name}")
