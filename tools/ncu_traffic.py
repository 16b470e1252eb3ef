"""ncu -i X.ncu-rep --page raw --csv | python tools/ncu_traffic.py -> {kernel: dram bytes per launch} for the first
launch of every kernel name in the capture (bench.py reads profiles/traffic.json for roofline.traffic)"""
import csv, json, subprocess, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}; units = rows[1]
def val(r, k):
    v = float(r[ix[k]].replace(',', ''))
    u = units[ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {}
big = {}
for r in rows[2:]:
    name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').split('<')[0].strip()
    grid = r[ix['Grid Size']]
    key = name
    g = int(grid.strip("()").split(",")[0])
    if key in out and g <= big.get(key, 0):   # the level-1 launch = the largest grid of that kernel
        continue
    big[key] = g
    out[key] = int(val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum'))
    out[key + "_grid"] = grid
try:
    out["commit"] = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    out["commit"] = "HEAD of the snapshot sent by gpurun"
out["source"] = "ncu --set full --clock-control none, tools/prof_fwdinv.py (C2, one 4096^2 image), first launch of each kernel"
print(json.dumps(out, indent=1))
