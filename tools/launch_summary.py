"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
   python tools/launch_summary.py launches.csv [first_id last_id]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ci = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr): continue
    try:
        rows.append((int(r[ci["ID"]]), r[ci["Kernel Name"]], float(r[ci["Metric Value"]]), r[ci["Metric Unit"]]))
    except ValueError:
        pass
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
def us(v, u): return v / 1000 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1000
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for i, name, v, u in rows:
    if not (lo <= i <= hi): continue
    n = re.sub(r"\(.*", "", name)
    n = re.sub(r"<.*", "", n)
    t = us(v, u)
    agg[n][0] += 1; agg[n][1] += t; tot += t
print(f"launches {sum(a[0] for a in agg.values())}, total {tot/1000:.2f} ms (ids {lo}..{hi})")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{t/1000:9.3f} ms {100*t/tot:5.1f}%  x{c:<5d} {n[:100]}")
