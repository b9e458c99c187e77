#!/usr/bin/env python
"""Top source lines of a kernel by executed warp instructions / stall samples, from an .ncu-rep captured with
--import-source on (kernels compiled with -lineinfo).  usage: ncu_source_hot.py rep [n_samples_total] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
denom = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:]))
    try: out.append((cur, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]), d))
    except Exception: pass
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total warp instr {ti}  ({ti/denom:.2f} per unit), stall samples {ts}")
print("== by instructions")
for f, ln, src, n, s, d in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{n/denom:7.2f} {100*n/ti:5.1f}%  st {100*s/ts:5.1f}%  {f}:{ln:<4d} {src}")
print("== by stall samples")
for f, ln, src, n, s, d in sorted(out, key=lambda o: -o[4])[:25]:
    reasons = sorted(((int(d[k]), k) for k in d if k.startswith("stall_") and "Not Issued" not in k and d[k].isdigit()), reverse=True)[:3]
    print(f"st {100*s/ts:5.1f}%  instr {n/denom:6.2f}  {f}:{ln:<4d} {src[:70]}  {[(k[6:],v) for v,k in reasons]}")
