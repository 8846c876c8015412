"""ncu raw page (csv) -> the short `metric,unit,value` summaries kept under profiles/ (same metric selection as the earlier ones).

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv profiles/r02_k_eval_8192_128regs_ncu_full_summary.csv > profiles/<new>_ncu_full_summary.csv
"""
import csv
import sys

raw, like = sys.argv[1], sys.argv[2]
want = [row[0] for row in csv.reader(open(like)) if row][1:]
rows = list(csv.reader(open(raw)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
col = {n: i for i, n in enumerate(names)}
out = csv.writer(sys.stdout)
out.writerow(["metric", "unit", "value"])
for m in want:
    if m in col:
        out.writerow([m, units[col[m]], vals[col[m]]])
