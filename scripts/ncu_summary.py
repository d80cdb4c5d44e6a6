"""Print the key raw metrics of every kernel in an ncu report. usage: ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:80])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:72s} {r[i]} {units[i]}")
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and r[i]]
    print('  stalls:', ', '.join(f"{h.split('issue_stalled_')[1].split('_per')[0]}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
