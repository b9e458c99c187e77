"""Print the kernels of the last step from an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: python tools/launch_list.py gpurun_out/train_launches.csv [first-kernel-substring]"""
import csv
import sys

path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 else "fold_tc"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")))
        for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
starts = [i for i, r in enumerate(rows) if anchor in r[1]]
step = rows[starts[-1]:] if starts else rows
total = sum(r[2] for r in step)
for i, name, ns in step:
    print(f"{i:5d} {ns / 1e3:9.1f} us  {100 * ns / total:5.1f}%  {name[:100]}")
print(f"total {total / 1e3:.1f} us over {len(step)} launches")
