"""Top stall locations of an `ncu --page source --csv` dump (SASS view): usage ncu_top_stalls.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        v = float(r[ci["# Samples"]])
    except Exception:
        continue
    data.append((v, idx, r))
tot = sum(v for v, _, _ in data)
print("total samples", tot)
for v, idx, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((float(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{int(v):7d} {100*v/tot:5.1f}%  #{idx:5d} {r[ci['Source']].strip()[:70]:70s} {top[0][1]}={int(top[0][0])} {top[1][1]}={int(top[1][0])}")
