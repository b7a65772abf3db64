"""Aggregate an `ncu --csv --metrics ...` launch list per kernel (time share + mean metrics)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        v = v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(r[ui], 1e6)   # -> ms
    per.setdefault(r[ii], {"name": r[ki].split("(")[0][:48]})[r[mi]] = v
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for d in per.values():
    cnt[d["name"]] += 1
    for k, v in d.items():
        if k != "name":
            agg[d["name"]][k] += v
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print("total GPU time %.3f ms over %d launches" % (tot, len(per)))
for name, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    n = cnt[name]
    t = a["gpu__time_duration.sum"]
    extra = "  ".join("%s=%.3g" % (k.split(".")[0].replace("smsp__", "").replace("sm__", ""), (v if k.endswith(".sum") else v / n))
                      for k, v in a.items() if k != "gpu__time_duration.sum")
    print("%-48s n=%4d  %9.3f ms %5.1f%%  avg %.4f ms | %s" % (name, n, t, 100 * t / tot, t / n, extra))
