"""One line per profiled launch from an `ncu --page raw --csv` dump: the counters the DESIGN tables quote."""
import csv, sys
COLS = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'dram_rd_MB', 1e-6), ('dram__bytes_write.sum', 'dram_wr_MB', 1e-6),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1), ('lts__t_bytes.sum', 'L2_MB', 1e-6),
        ('l1tex__t_sector_hit_rate.pct', 'L1hit%', 1), ('lts__t_sector_hit_rate.pct', 'L2hit%', 1),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst', 1), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 1),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 1), ('smsp__inst_executed.sum', 'Minst', 1e-6),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_longsb', 1),
        ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'stall_membar', 1),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier', 1),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall_lg', 1),
        ('launch__grid_size', 'grid', 1), ('launch__block_size', 'block', 1), ('launch__registers_per_thread', 'regs', 1)]
rows = list(csv.reader(open(sys.argv[1])))
while rows and (not rows[0] or rows[0][0] != 'ID'):
    rows.pop(0)
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
for r in rows[2:]:
    name = r[idx['Kernel Name']].replace('void ', '').replace('vrad::', '')
    name = name.split('(')[0][:44]
    out = []
    for c, label, scale in COLS:
        if c not in idx:
            continue
        try:
            x = float(r[idx[c]].replace(',', ''))
            u = units[idx[c]]
            if label == 'us' and u in ('us', 'usecond'): x *= 1e3
            if label == 'us' and u in ('ms', 'msecond'): x *= 1e6
            if label.endswith('MB') and u == 'Kbyte': x *= 1e3
            if label.endswith('MB') and u == 'Mbyte': x *= 1e6
            if label.endswith('MB') and u == 'Gbyte': x *= 1e9
            out.append(f"{label}={x * scale:.1f}")
        except ValueError:
            pass
    print(name.ljust(44), ' '.join(out))
