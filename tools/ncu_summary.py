#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): key metrics per kernel, stall
reasons and opcode mix from the SASS page. Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    print(d['Kernel Name'][:140])
    for w in WANT:
        if w in d:
            print('  %-70s %s %s' % (w, d[w], units[hdr.index(w)]))
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
kern, h2, data = None, None, collections.OrderedDict()
for r in csv.reader(io.StringIO(sass)):
    if r and r[0] == 'Kernel Name':
        kern = r[1][:70]; data[kern] = []; h2 = None; continue
    if r and r[0] == 'Address':
        h2 = r; continue
    if kern and h2 and len(r) == len(h2):
        data[kern].append(dict(zip(h2, r)))
for k, v in data.items():
    tot = sum(int(x['Instructions Executed']) for x in v)
    print('-' * 100)
    print(k, '| SASS lines', len(v), '| warp instructions', tot)
    agg = collections.Counter()
    for x in v:
        for key in x:
            if key.startswith('stall_') and 'Not Issued' not in key:
                agg[key] += int(x[key] or 0)
    st = sum(agg.values()) or 1
    print('  stall samples:', ', '.join('%s %.0f%%' % (a[6:], 100.0 * b / st) for a, b in agg.most_common(9)))
    op = collections.Counter()
    for x in v:
        s = x['Source'].strip()
        m = s.split()[0] if not s.startswith('@') else s.split()[1]
        op[m.split('.')[0]] += int(x['Instructions Executed'])
    print('  opcode mix   :', ', '.join('%s %.0f%%' % (a, 100.0 * b / tot) for a, b in op.most_common(16)))
