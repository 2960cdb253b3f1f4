"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel (and per grid shape)."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0.0, 0])
per_shape = collections.defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = name.replace("(anonymous namespace)::", "")
    tot[name][0] += us; tot[name][1] += 1
    per_shape[(name, r["Grid Size"], r["Block Size"])][0] += us; per_shape[(name, r["Grid Size"], r["Block Size"])][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total/1e3:.2f} ms over {sum(v[1] for v in tot.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]/1e3:9.3f} ms {100*v[0]/total:5.1f}%  n={v[1]:5d}  avg {v[0]/v[1]:8.1f} us  {k}")
if len(sys.argv) > 2:
    print("--- top shapes")
    for k, v in sorted(per_shape.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2])]:
        print(f"{v[0]/1e3:9.3f} ms n={v[1]:4d} avg {v[0]/v[1]:8.1f} us  {k}")
