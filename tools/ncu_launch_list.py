"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, share.
usage: python tools/ncu_launch_list.py launches.csv "header comment" > profiles/xx_launches.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name, unit, val = r[4], r[-2], float(r[-1].replace(",", ""))
    ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    tot[name] += ns; cnt[name] += 1
total = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"# per kernel: launches, total ms, share of the profiled device time {total / 1e6:.1f} ms ({len(rows)} launches; cold-cache, serialised)")
for name, ns in tot.most_common():
    print(f"{cnt[name]:5d} {ns / 1e6:10.3f} ms {100 * ns / total:6.2f} %  {name[:110]}")
