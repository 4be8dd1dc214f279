"""ncu launch list (gpu__time_duration.sum CSV) -> per-kernel markdown summary for profiles/."""
import collections, csv, re, sys
src, title = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0}.get(row["Metric Unit"], 1.0)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"# {title}\n\nncu `--metrics gpu__time_duration.sum --clock-control none` over ONE config-2 clip (cold-cache, serialised launches: compare shares).\n")
print(f"total kernel time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| `{k[:70]}` | {n} | {ms:.2f} | {100*ms/tot:.1f}% | {1000*ms/n:.1f} |")
