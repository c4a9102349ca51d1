"""Hot spots of an `ncu --page source --csv` export: samples per SASS region + top instructions.
usage: python tools/sass_hot.py file.csv [bin]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
binw = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = rows[1]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= iX: continue
    try: data.append((r[iS].strip(), int(r[iN]), int(r[iX]), r))
    except ValueError: pass
tot = sum(d[1] for d in data); totx = sum(d[2] for d in data)
print(f"{len(data)} SASS instructions, {tot} samples, {totx} warp-instructions executed")
for b in range(0, len(data), binw):
    seg = data[b:b + binw]
    s = sum(d[1] for d in seg); x = sum(d[2] for d in seg)
    st = {}
    for d in seg:
        for i in stall_cols:
            v = int(d[3][i] or 0)
            if v: st[hdr[i]] = st.get(hdr[i], 0) + v
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"[{b:5d}-{b + len(seg):5d}] samples {s:6d} ({100 * s / tot:5.1f}%) exec {x:8d} ({100 * x / totx:5.1f}%)  {top}   | {seg[0][0][:40]}")
print("top instructions:")
for i, d in sorted(enumerate(data), key=lambda kv: -kv[1][1])[:25]:
    print(f"  #{i:5d} samples {d[1]:5d} exec {d[2]:7d}  {d[0][:90]}")
