"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) -> per-kernel table (launches, total ms, share, avg us).
  python scripts/launches_summarize.py gpurun_out/r02_launches.csv profiles/r02_launches.md"""
import collections
import csv
import re
import sys


def main(src, dst):
    lines = [l for l in open(src, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    head = rows[0]
    ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("vg::"))
    with open(dst, "w") as f:
        f.write(f"total kernel time {total / 1e3:.1f} ms over {n} launches; `vg::` kernels {ours / total:.1%} of it\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e3:.2f} | {a[1] / total:.1%} | {a[1] / a[0]:.1f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
