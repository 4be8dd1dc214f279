"""Per-kernel SASS mnemonic counts of libvglamm.so (cuobjdump -sass): which kernels use the 5th-gen tensor cores (UTCHMMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor loads/stores/reductions, UBLKCP = bulk copies) and which still run the legacy
warp-level HMMA (mma.sync). Runs without a GPU:   python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "videoglamm_b200" / "lib" / "libvglamm.so"
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "SYNCS", "MUFU", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts, regs, cur = collections.OrderedDict(), {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
            counts[cur]["total"] += 1
    print("# SASS summary of `videoglamm_b200/lib/libvglamm.so` (cuobjdump -sass, sm_100a), round 2\n")
    print("Counts of static instructions per kernel. UTCHMMA = `tcgen05.mma`, LDTM/STTM = `tcgen05.ld/st`, UTMALDG/UTMASTG/UTMAREDG = TMA tensor "
          "load / store / reduce, UBLKCP = `cp.async.bulk`, HMMA = legacy `mma.sync`. Regenerate: `python scripts/sass_summary.py`.\n")
    print("| kernel | instr | " + " | ".join(OPS) + " |")
    print("|---|---:|" + "---:|" * len(OPS))
    for k, c in sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["HMMA"], kv[0])):
        print(f"| `{k}` | {c['total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    tc = [k for k, c in counts.items() if c["UTCHMMA"]]
    hm = [k for k, c in counts.items() if c["HMMA"] and not c["UTCHMMA"]]
    print(f"\ntcgen05 kernels: {len(tc)}; legacy HMMA-only kernels: {len(hm)} ({', '.join('`' + k + '`' for k in hm)}); total kernels: {len(counts)}")


if __name__ == "__main__":
    sys.exit(main())
