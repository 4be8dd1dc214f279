"""ncu raw CSV (`ncu -i X.ncu-rep --page raw --csv`) -> compact per-launch CSV + markdown table for profiles/.
  python scripts/ncu_summarize.py gpurun_out/r02_targets_raw.csv profiles/r02_ncu_targets"""
import csv
import sys

COLS = {"us": "gpu__time_duration.sum", "dram_read_MB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
        "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_mem_pct": "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "hmma_pct": "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "smem_KB": "launch__shared_mem_per_block_dynamic"}
UNIT = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {}
    for k, name in COLS.items():
        cands = [i for i, c in enumerate(head) if c == name or c.endswith("." + name)]
        if not cands:
            cands = [i for i, c in enumerate(head) if name in c]
        idx[k] = cands[0] if cands else None
    ki, gi, bi = head.index("Kernel Name"), head.index("Grid Size"), head.index("Block Size")
    out = []
    for r in data:
        d = {"id": r[0], "kernel": r[ki].split("(")[0].replace("void ", ""), "grid": r[gi], "block": r[bi]}
        for k, i in idx.items():
            if i is None or r[i] in ("", "no data", "n/a"):
                d[k] = ""
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if k.endswith("_MB") or k == "us":
                v *= UNIT.get(u, 1.0)
            if k == "smem_KB":
                v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u.split("/")[0], 1.0)
            d[k] = round(v, 3)
        out.append(d)
    keys = ["id", "kernel", "grid", "block"] + list(COLS)
    with open(dst + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(out)
    with open(dst + ".md", "w") as f:
        f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for d in out:
            f.write("| " + " | ".join(str(d[k]) for k in keys) + " |\n")
    print(open(dst + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
