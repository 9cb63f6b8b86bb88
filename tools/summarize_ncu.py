"""Turns the ncu captures brought back in gpurun_out/ into the committed summaries under profiles/<round>/:
  ncu_full_summary.json (selected --set full metrics per kernel), profiles/traffic.json (DRAM bytes per unit, read by
  bench.py) and launches_bench_summary.txt (per-kernel share of the bench command's launch list).
usage: python tools/summarize_ncu.py <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <launches.csv> <out_dir> <profiled_batch>"""
import collections, csv, json, sys

raw_csv, launches_csv, out_dir, B = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(raw_csv)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('math_pipe_throttle', 'long_scoreboard', 'wait', 'barrier', 'short_scoreboard', 'not_selected', 'dispatch_stall', 'mio_throttle')]


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return s


out = {}
for r in data:
    name = r[idx['Kernel Name']]
    if 'ntt_kernel' in name:
        short = 'ntt_kernel_inv' if name.split('(')[0].rstrip().endswith('1>') or '(bool)1' in name else 'ntt_kernel_fwd'
    else:
        short = 'ks_fused_kernel_mul_relin'
    out[short] = {'kernel': name[:140], **{w: {'value': num(r[idx[w]]), 'unit': units[idx[w]]} for w in WANT if w in idx}}
json.dump(out, open(out_dir + '/ncu_full_summary.json', 'w'), indent=1)

SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
traffic = {}
for k, d in out.items():
    rd = d['dram__bytes_read.sum']['value'] * SCALE.get(d['dram__bytes_read.sum']['unit'], 1)
    wr = d['dram__bytes_write.sum']['value'] * SCALE.get(d['dram__bytes_write.sum']['unit'], 1)
    units_per_launch = B * 2 * 4 if 'ntt' in k else B
    alg = units_per_launch * (131072 if 'ntt' in k else 1572864)
    traffic[k] = {'dram_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr, 'algorithmic_bytes_per_launch': alg,
                  'ratio': (rd + wr) / alg, 'profiled_batch': B, 'dram_bytes_per_unit': (rd + wr) / units_per_launch,
                  'warp_instructions_per_unit': d['smsp__inst_executed.sum']['value'] / units_per_launch}
    print(k, 'dram/alg %.3f' % traffic[k]['ratio'], 'inst/unit %.0f' % traffic[k]['warp_instructions_per_unit'],
          'issue %.1f%%' % d['smsp__issue_active.avg.pct_of_peak_sustained_active']['value'])
json.dump(traffic, open('profiles/traffic.json', 'w'), indent=1)

rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    us = float(r[i_val].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}.get(r[i_unit], 1)
    k = r[i_name].split('(')[0][:70]
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
with open(out_dir + '/launches_bench_summary.txt', 'w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 60: python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = "%-72s launches=%3d total=%10.1f us share=%5.1f%%" % (k, n, us, 100 * us / tot)
        print(line)
        f.write(line + "\n")
