"""Generic: selected metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump -> JSON.
usage: python tools/summarize_ncu_raw.py <raw.csv> <out.json>"""
import csv, json, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('math_pipe_throttle', 'long_scoreboard', 'wait', 'barrier', 'short_scoreboard', 'not_selected', 'dispatch_stall', 'mio_throttle')]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in data:
    d = {"kernel": r[idx['Kernel Name']][:120]}
    for w in WANT:
        if w in idx:
            try:
                v = float(r[idx[w]].replace(',', ''))
            except ValueError:
                v = r[idx[w]]
            d[w] = {"value": v, "unit": units[idx[w]]}
    out.append(d)
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print("wrote", sys.argv[2], len(out), "kernels")
