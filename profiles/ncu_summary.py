#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (ncu -i ... --page raw --csv) into the handful of numbers
DESIGN.md / bench.py quote.  Usage: ncu_summary.py file.ncu-rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units, vals = r[0], r[1], r[2 + row]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
def g(k):
    v = d.get(k, ('', ''))
    return f'{v[0]} {v[1]}'.strip()
print('kernel:', g('Kernel Name')[:80], '| grid', g('launch__grid_size'), 'block', g('launch__block_size'))
for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
          'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__warps_eligible.avg.per_cycle_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
          'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']:
    print(f'  {k:70s} {g(k)}')
st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(v[0].replace(',', '') or 0)
      for h, v in d.items() if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('not_issued')}
tot = sum(st.values()) or 1
print('  warp-state samples:', ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
