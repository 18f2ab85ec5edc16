"""Group the SASS of an `ncu --page source --csv` dump into contiguous regions of equal execution count and print each
region's share of all warp instructions with its opcode mix.  usage: ncu_regions.py src.csv [min_share]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
col = {h: i for i, h in enumerate(rows[1])}
data = rows[2:]
tot = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
regions = []
cur = None
for r in data:
    n = int(r[col["Instructions Executed"]] or 0)
    op = r[col["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0] if op else "?"
    if cur is None or abs(n - cur["n"]) > 0.02 * max(n, cur["n"], 1):
        cur = {"n": n, "cnt": 0, "ops": collections.Counter(), "start": r[col["Address"]][-5:], "smp": 0}
        regions.append(cur)
    cur["cnt"] += 1; cur["ops"][op] += 1; cur["smp"] += int(r[col["# Samples"]] or 0)
print(f"total warp instr {tot}")
for g in regions:
    share = g["n"] * g["cnt"] / tot
    if share >= thr:
        print(f'{g["start"]} {100*share:5.1f}%  {g["cnt"]:4d} instr x {g["n"]:>11d}  smp={g["smp"]:>7d}  ' +
              " ".join(f"{k}:{v}" for k, v in g["ops"].most_common(9)))
