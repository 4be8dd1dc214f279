"""Per-instruction stall summary from `ncu --page source --csv` output (possibly several kernels concatenated)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for bi, b in enumerate(blocks):
    print(bi, b["name"][:100], len(b["rows"]))
b = blocks[which]
hdr = b["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print("total samples", tot, sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    st = sorted(((s, int(r[idx[s]])) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(r[idx['# Samples']].rjust(6), r[idx['Instructions Executed']].rjust(8), r[idx['Source']][:64].ljust(64), st)
