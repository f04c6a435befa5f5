"""Summarise an .ncu-rep (raw page) into a small CSV for profiles/: python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv, subprocess, sys, io
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_static', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__sass_average_branch_targets_threads_uniform.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform.sum', 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum', 'launch__occupancy_limit_blocks', 'sm__ctas_launched.sum']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], 'w') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rows) - 2)])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
            print(k, '=', [r[i] for r in rows[2:]], units[i])
