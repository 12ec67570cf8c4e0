"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/ncu_table.py launches.csv [first_launch [n_launches]]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            out.append((d["Kernel Name"].split("(")[0][:48], float(d["Metric Value"].replace(",", "")) / 1e3))
        except ValueError:
            pass
out = out[lo:lo + cnt]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in out:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s n=%4d total=%9.1f us avg=%8.1f us %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print("total %.1f us over %d launches" % (tot, len(out)))
