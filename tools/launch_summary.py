"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py <launches.csv> <out.json> "<command>" """
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0][:80]
    a = agg.setdefault(name, {"launches": 0, "ms": 0.0})
    a["launches"] += 1
    a["ms"] += float(r[mv].replace(",", "")) / 1e6
tot = sum(a["ms"] for a in agg.values())
for a in agg.values():
    a["share"] = a["ms"] / tot
json.dump({"command": sys.argv[3], "total_ms": tot, "kernels": agg}, open(sys.argv[2], "w"), indent=1)
top = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:4]
print(tot, [(k, round(v["share"], 4)) for k, v in top])
