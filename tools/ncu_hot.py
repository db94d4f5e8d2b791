"""Top stall-sample instructions from an `ncu --page source --csv` export.  python tools/ncu_hot.py file.csv [N]"""
import csv, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # which profiled launch inside the file
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[sec]; end = his[sec + 1] - 1 if sec + 1 < len(his) else len(rows)
print(rows[hi - 1][1][:120] if hi > 0 else "")
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("total samples", tot, "instructions", len(data))
print("stall mix:", ", ".join(f"{k[6:]}={v*100//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {s*100/max(tot,1):5.1f}%  {r[ix['Source']].strip()[:90]:90s} {top}")
