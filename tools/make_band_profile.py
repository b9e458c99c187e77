#!/usr/bin/env python
"""profiles/r02_band_kernel.md (+ kernel_metrics.json, traffic.json entries) from an ncu --set full capture of the band
kernel on the bench command and the launch list of the same command.
usage: python tools/make_band_profile.py <render_band_full.ncu-rep> <launches.csv> [notes.md]"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches = sys.argv[1], sys.argv[2]
notes = open(sys.argv[3]).read() if len(sys.argv) > 3 else ""
SAMPLES = 245760 * 64
src = open(os.path.join(ROOT, "behindthescenes_b200/csrc/render_band.cu")).read().split("\n")
marks = [("helpers (fp16 splits, packing)", r"^namespace band \{"), ("sincos", r"void sincos_moderate"), ("far-sample code (out of line)", r"void code_far_to_tmem"),
         ("one-time setup + camera lambda", r"^template <int HD, bool kExact, int VC>"), ("per-pass setup", r"---- per-pass setup"), ("P0 depth, projection, taps", r"// P0 of a chunk"),
         ("band geometry + bulk copies", r"// footprints of the 4 warps"), ("P1a positional code -> TMEM", r"// P1a:"), ("P1b weight scatter", r"// P1b:"),
         ("prologue of a pass", r"---- prologue: operands"), ("MMA issue", r"MMA issue \(one thread"), ("next-tile glue + weight clearing", r"P0 of the next chunk, while"),
         ("P3 epilogue", r"P3: accumulator -> sigma"), ("C compositing", r"---- C: alpha compositing"), ("D colour", r"---- D: colour views"), ("per-ray outputs", r"---- per-ray outputs"),
         ("pack_g16 / launchers", r"^// G \(n, H, W, HD\) fp32 -> operand blocks")]
pos = []
for name, pat in marks:
    for i, l in enumerate(src):
        if re.search(pat, l):
            pos.append((i + 1, name)); break
pos.sort()
ranges = [(lo, (pos[k + 1][0] - 1 if k + 1 < len(pos) else len(src)), name) for k, (lo, name) in enumerate(pos)]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = hdr = None
acc, other, lines = {}, {}, []
ti = ts = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:]))
    try: ln, n, s = int(r[0]), int(d["Instructions Executed"]), int(d["# Samples"])
    except Exception: continue
    ti += n; ts += s
    lines.append((cur, ln, r[1].strip()[:80], n, s, d))
    if cur == "render_band.cu":
        for lo, hi, name in ranges:
            if lo <= ln <= hi:
                a = acc.setdefault(name, [0, 0]); a[0] += n; a[1] += s; break
    else:
        a = other.setdefault(cur, [0, 0]); a[0] += n; a[1] += s
phase = [f"total {ti / SAMPLES:.2f} warp-instructions per sample, {ts} stall samples"]
for name, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    phase.append(f"{n / SAMPLES:7.2f} {100 * n / ti:5.1f}%  stalls {100 * s / ts:5.1f}%  {name}")
for name, (n, s) in sorted(other.items(), key=lambda kv: -kv[1][0])[:5]:
    phase.append(f"{n / SAMPLES:7.2f} {100 * n / ti:5.1f}%  stalls {100 * s / ts:5.1f}%  [file] {name}")
stall_lines = []
sraw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sraw)))
sh = next(i for i, r in enumerate(srows) if "Source" in r and "# Samples" in r)
shdr, sdata = srows[sh], srows[sh + 1:]
iv = lambda x: int(x) if x.isdigit() else 0
c_s, c_src, c_e = shdr.index("# Samples"), shdr.index("Source"), shdr.index("Instructions Executed")
reasons = [h for h in shdr if h.startswith("stall_") and "Not" not in h]
stot = sum(iv(r[c_s]) for r in sdata)
for j in sorted(range(len(sdata)), key=lambda j: -iv(sdata[j][c_s]))[:8]:
    r = sdata[j]
    why = sorted(((iv(r[shdr.index(h)]), h[6:]) for h in reasons), reverse=True)[:2]
    prev = sdata[j - 1][c_src].strip()[:48] if j else ""
    stall_lines.append(f"{100 * iv(r[c_s]) / stot:5.1f}%  {r[c_src].strip()[:52]:52s} (after {prev})  {[(k, v) for v, k in why if v]}")
rawm = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(rawm)))
m = dict(zip(rr[0], rr[2] if len(rr) > 2 else rr[1]))
g = lambda k: float(m[k].replace(",", ""))
mix = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v) for k, v in m.items()
       if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(v) > 0.01}
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools/ncu_summary.py"), rep], capture_output=True, text=True).stdout
ll = subprocess.run([sys.executable, os.path.join(ROOT, "tools/launch_list.py"), launches], capture_output=True, text=True).stdout
units = dict(zip(rr[0], rr[1]))
scale = lambda k: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get(k, "byte"), 1.0)
dram = sum(g(k) * scale(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
issue, lsu, tens = g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
doc = f"""# Round 2: band render kernel (`render_band_kernel<64, exact, 1 view>`), bench cfg2 stereo frame, 1x B200

Commands (under gpurun, one GPU):

* launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-extras` (-> `profiles/r02_launches.csv`)
* full capture: `ncu --set full --clock-control none --import-source on -k regex:render_band -s 12 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-extras`

## Last step of the launch list (cold-cache, serialised: shares, not absolutes)
```
{ll}```
The render kernel is ~91 % of the step (bench, live CUDA events: 0.91 ms of 0.99 ms); the work-list pass of the gather kernel finds
nothing to do on this frame; the per-frame glue (fold, G16 packing of three copies, colour packing) is 6 %.

{summ}
## Where the instructions go (warp-instructions per sample, {SAMPLES / 1e6:.2f} M samples per launch)
```
{chr(10).join(phase)}
```
(`tcgen05.cuh` = barrier waits + MMA / TMEM wrappers; helpers + `cuda_fp16.hpp` = the fp16 hi/lo operand splits; `common.cuh` =
projection, depth code, taps, jitter hash.)  The r01 gather kernel executed 59.7 warp-instructions per sample; its bilinear gather
of G (12.3 per sample, a third of its stalls) is tensor-core work here, and the local-memory spills are gone.

## What binds
Issue slots {issue:.1f} % busy at 16 resident warps per SM (registers: 128 x 512 threads; shared memory: 2 x 114 KB; TMEM: 2 x 256
columns -- every one of the three caps residency at 4 tile groups per SM); tensor pipe {tens:.1f} %; LSU data pipe {lsu:.1f} %
(mostly shared memory: operand reads of the MMAs, the weight scatter with its 4-way bank conflicts, w_out broadcasts); DRAM
{dram / 1e6:.0f} MB per launch against 266 MB algorithmic.  Warp states per issued instruction:
```
{json.dumps(mix, indent=1)}
```
Top stall sites (SASS, share of all warp-state samples; the two SYNCS waits are the accumulator-ready and footprint barriers, BAR the tile-group barrier):
```
{chr(10).join(stall_lines)}
```
{notes}"""
open(os.path.join(ROOT, "profiles/r02_band_kernel.md"), "w").write(doc)
kmp = os.path.join(ROOT, "profiles/kernel_metrics.json")
km = json.load(open(kmp)) if os.path.isfile(kmp) else {}
km["band"] = {"issue_slots_busy_pct": issue, "lsu_data_pipe_pct": lsu, "tensor_pipe_pct": tens, "warp_instr_per_sample": round(ti / SAMPLES, 2),
              "source": "profiles/r02_band_kernel.md (ncu --set full of the bench command)"}
json.dump(km, open(kmp, "w"), indent=1)
tp = os.path.join(ROOT, "profiles/traffic.json")
t = json.load(open(tp))
t["band"] = {"dram_bytes_per_launch": dram, "source": "profiles/r02_band_kernel.md"}
json.dump(t, open(tp, "w"), indent=1)
print(doc[:600]); print(km["band"], t["band"])
