"""ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py  -> one line of key metrics per profiled launch"""
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
def g(r, k, d=1.0):
    try: return float(r[ix[k]].replace(',', '')) / d
    except Exception: return float('nan')
keys = [("dur_us", "gpu__time_duration.sum", 1), ("dram_rd_MB", "dram__bytes_read.sum", 1), ("dram_wr_MB", "dram__bytes_write.sum", 1),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("inst_M", "smsp__inst_executed.sum", 1e6),
        ("lsu_shared_wavefronts_M", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1e6),
        ("smem_conflicts_M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1e6),
        ("l2_MB", "lts__t_bytes.sum", 1)]
units = rows[1]
for r in rows[2:]:
    name = r[ix['Kernel Name']].split('(')[0]
    out = [f"{r[ix['ID']]:>2} {name[:34]:34s} grid={r[ix['Grid Size']]}"]
    for lab, k, d in keys:
        v = g(r, k, d)
        u = units[ix[k]] if k in ix else ''
        if lab.endswith('MB') and u == 'Gbyte': v *= 1000
        if lab.endswith('MB') and u == 'Kbyte': v /= 1000
        if lab.endswith('MB') and u == 'byte': v /= 1e6
        out.append(f"{lab}={v:.2f}")
    stalls = {k.split('issue_stalled_')[1].split('_per_')[0]: g(r, k) for k in ix if 'smsp__average_warps_issue_stalled' in k and k.endswith('.ratio')}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
    out.append("stalls/issue: " + ", ".join(f"{k}={v:.2f}" for k, v in top))
    print("  ".join(out))
