"""Print the headline metrics of an .ncu-rep (raw page) -- what DESIGN.md / profiles/*.md quote.

    python profiles/ncu_keys.py gpurun_out/k3.ncu-rep
"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'smsp__cycles_elapsed.avg.per_second']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    print('==', rep, '|', vals[hdr.index('Kernel Name')][:90])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'  {k:64s} {vals[i]:>18s} {units[i]}')
