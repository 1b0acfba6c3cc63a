"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: the last `n` launches of this library's kernels."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
items = [(int(r[ii]), r[ki][:70], float(r[vi].replace(",", ""))) for r in rows[1:]]
ours = [x for x in items if any(t in x[1] for t in ("ern::", "simtc::", "small::", "gemmtc::", "combiner::", "visualsr::", "dvr::", "bbcloss::"))]
for x in ours[-n:]:
    print(x)
print("sum_us", sum(x[2] for x in ours[-n:]) / 1e3)
