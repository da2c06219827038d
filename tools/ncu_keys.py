"""Print the metrics that matter from an `ncu --page raw --csv` dump (development tool)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_op_red.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__m_xbar2l1tex_read_sectors.sum',
        'l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_set_conflicts_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or None
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print('----', name[:110])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:78s} {r[i]:>18s} {units[i]}")
    if want:
        for k in hdr:
            if any(w in k for w in want):
                i = hdr.index(k)
                print(f"  {k:78s} {r[i]:>18s} {units[i]}")
