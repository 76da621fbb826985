"""Summarise an `ncu --set full` raw CSV export (ncu -i X.ncu-rep --page raw --csv) into profiles/<name>.csv."""
import csv
import sys

raw, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
want += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
want += [h for h in hdr if h.startswith('smsp__sass_thread_inst_executed_op_d') and h.endswith('.sum')]
with open(out, 'w') as f:
    f.write('# %s\nmetric,unit,value\n' % note)
    for k in want:
        for i, h in enumerate(hdr):
            if h == k:
                f.write('%s,%s,%s\n' % (k, units[i], vals[i]))
print(open(out).read())
