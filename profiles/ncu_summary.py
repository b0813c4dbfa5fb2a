"""Print the key metrics of an .ncu-rep (raw page) per captured kernel.  usage: python profiles/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__inst_executed_pipe_fp64.sum','sm__cycles_elapsed.max',
        # L2 / atomic (red) traffic: sectors of red.global.add reaching L2, their L2 lookup hits / misses, L2 throughput
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_red.sum.per_second',
        'lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum',
        'l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum', 'l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum',
        'lts__t_sector_op_read_hit_rate.pct', 'lts__t_sector_op_write_hit_rate.pct']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:80])
    for w in WANT:
        if w in hdr:
            print('  %-85s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
