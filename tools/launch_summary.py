"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[r[ki][:64]][0] += 1
    agg[r[ki][:64]][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, serialised kernel time %.2f ms" % (sum(v[0] for v in agg.values()), tot / 1e6))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s n=%5d total=%9.3f ms avg=%9.1f us share=%.3f" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, v[1] / tot))
