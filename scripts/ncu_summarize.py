"""ncu raw CSV (`ncu -i X.ncu-rep --page raw --csv`) -> compact per-launch CSV + markdown table for profiles/.
  python scripts/ncu_summarize.py gpurun_out/r02_targets_raw.csv profiles/r02_ncu_targets"""
import csv
import sys

COLS = {"us": "gpu__time_duration.sum", "dram_read_MB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
        "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_mem_pct": "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "hmma_pct": "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "smem_KB": "launch__shared_mem_per_block_dynamic"}
UNIT = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


WHAT = ["gemm Hiera-s3 fc1 32768x2304x576 bias+GELU", "gemm Hiera-s3 fc2 32768x576x2304 +resid fp32", "gemm Phi-3 prefill gate_up 1696x16384x3072 SwiGLU",
        "gemm Phi-3 prefill down 1696x3072x8192 +resid", "gemv Phi-3 decode gate_up 1x16384x3072 RMS+SwiGLU", "gemv Phi-3 decode down 1x3072x8192",
        "attn Phi-3 prefill 1696 causal 32x96", "attn SAM2 memory 4096x28700 1x256 split-K", "attn merge (split-K)", "attn Hiera global 8x4096x4096 8x72",
        "attn Hiera stage-3 windows 128x256 8x72", "attn Hiera stage-2 windows 8192x16 4x72", "attn Phi-3 decode 1x1709 32x96",
        "gemv batched decode gate_up 4x16384x3072 RMS+SwiGLU", "gemv batched decode down 4x3072x8192 +resid",
        "fused SAM2 mask decoder HW=4096 B=1 Nt=9, split-bf16 tensor-core contractions (default)", "fused SAM2 mask decoder, fp32-FMA contractions (force_exact)"]


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
        d = {"id": r[0], "what": WHAT[len(out)] if len(sys.argv) > 3 and len(out) < len(WHAT) else "", "kernel": r[ki].split("(")[0].replace("void ", "").replace("vg::", ""),
             "grid": r[gi], "block": r[bi]}
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
    keys = ["id", "what", "kernel", "grid", "block"] + list(COLS)
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
