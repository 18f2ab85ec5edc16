"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions (by warp-level executions) in address order.
usage: python tools/ncu_hot.py src.csv [min_fraction]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
print("total warp instr", tot, "n sass", len(data))
for r in data:
    n = int(r[col["Instructions Executed"]] or 0)
    if n >= thr * tot:
        print(f'{r[col["Address"]][-5:]} {100*n/tot:5.2f}% smp={r[col["# Samples"]]:>6s} {r[col["Source"]][:110]}')
