"""ncu -i X.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > f.csv;
python tools/ncu_source_stalls.py f.csv  -> stall reasons, samples by opcode, hottest instructions"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {k: i for i, k in enumerate(hdr)}
S = '# Samples'
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, S) for r in data)
print("total samples", tot, "instructions", len(data), "warp-instr executed", sum(I(r, 'Instructions Executed') for r in data))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(I(r, h) for r in data) for h in st}
print("reasons:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
byop = {}
for r in data:
    t = r[ix['Source']].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    d = byop.setdefault(op, [0, 0]); d[0] += I(r, S); d[1] += I(r, 'Instructions Executed')
print("by opcode (samples, executed):", ", ".join(f"{op}={s}/{n}" for op, (s, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -I(r, S))[:n]:
    reasons = sorted(((h[6:], I(r, h)) for h in st), key=lambda kv: -kv[1])[:2]
    print(str(I(r, S)).rjust(5), str(I(r, 'Instructions Executed')).rjust(7), r[ix['Source']].strip()[:84].ljust(84), reasons)
