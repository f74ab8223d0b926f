"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`, with their dominant stall reasons and
a few neighbouring instructions for orientation.   usage: ncu_top_sass.py source.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:n_top]
for i in order:
    r = data[i]
    st = sorted(((num(r, s), s) for s in stalls), reverse=True)[:3]
    print(f"#{i:5d} {100 * num(r, '# Samples') / tot:5.1f}%  exec {int(num(r, 'Instructions Executed')):8d}  {r[ix['Source']][:70]:70s} " +
          " ".join(f"{s[6:]}={int(v)}" for v, s in st if v > 0))
