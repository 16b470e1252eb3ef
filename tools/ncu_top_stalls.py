"""ncu -i X.ncu-rep --page source --csv | python tools/ncu_top_stalls.py <kernel index> [n]  -> most-sampled SASS lines"""
import csv, sys
want = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks = []; cur = None
for row in csv.reader(sys.stdin):
    if row and row[0] == "Kernel Name":
        cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(row)
b = blocks[want]; h = b[0]; data = b[1:]
ix = {k: i for i, k in enumerate(h)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("kernel", want, "of", len(blocks), "total samples", tot)
agg = {}
for r in data:
    for k in h:
        if k.startswith('stall_') and '(' not in k:
            agg[k[6:]] = agg.get(k[6:], 0) + int(r[ix[k]] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:n]:
    st = {k[6:]: int(r[ix[k]] or 0) for k in h if k.startswith('stall_') and '(' not in k}
    st = {k: v for k, v in st.items() if v > 0}
    print(r[ix['Address']][-5:], r[ix['Source']][:66].ljust(66), r[ix['# Samples']], st)
