"""Trim a `ncu --page raw --csv` export (2000+ columns) to the metrics the roofline discussion uses.
    python tools/ncu_trim.py in.csv out.csv"""
import csv
import io
import sys

KEEP = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit', 'smsp__average_warps_issue_stalled', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg ',
        'sm__cycles_elapsed.max', 'sm__ops_path_tensor', 'sm__pipe_tensor_cycles_active', 'sm__pipe_tensor_subpipe', 'gpu__dram_throughput',
        'dram__cycles_active.avg', 'sm__inst_executed_pipe_tensor', 'umma', 'utc']

lines = open(sys.argv[1], errors='replace').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
hdr = rows[0]
idx = [i for i, h in enumerate(hdr) if any(h == k or (len(k) > 3 and k in h) for k in KEEP)]
# drop the per-stall columns that are ~0 everywhere and the .min/.max/.sum duplicates of pct metrics
idx = [i for i in idx if not any(s in hdr[i] for s in ('.min.', '.max.', '.sum.pct', '.sum.per_second', '.peak_sustained', '.min', 'Triage'))
       or hdr[i] in ('ID', 'Kernel Name')]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i < len(r) else '' for i in idx])
print(f'{sys.argv[2]}: {len(rows) - 2} kernels x {len(idx)} columns')
