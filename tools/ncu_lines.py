#!/usr/bin/env python
"""Per-source-line executed instructions of a capture, as instructions per thread per tile (= per sample).
usage: python tools/ncu_lines.py <rep> <file-substring> [min]"""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
mn = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
WARPS = 245760 * 64 / 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = hdr = None
tot = 0
for r in csv.reader(io.StringIO(raw)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or want not in (cur or ""): continue
    d = dict(zip(hdr, r))
    try: n = int(d["Instructions Executed"]); s = int(d["# Samples"])
    except Exception: continue
    tot += n
    if n / WARPS >= mn: print(f"{n / WARPS:7.1f} st {s:5d}  {r[0]:>4s} {r[1].strip()[:110]}")
print("total", tot / WARPS)
