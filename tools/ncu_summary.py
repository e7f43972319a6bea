#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X.summary.txt"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]}  (ncu --set full --clock-control none; per-launch, cold cache, serialised)")
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'  {w}: {r[i]} {units[i]}')
    try:
        rd = float(r[hdr.index('dram__bytes_read.sum')]); wr = float(r[hdr.index('dram__bytes_write.sum')]); t = float(r[hdr.index('gpu__time_duration.sum')])
        ur, ut = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('gpu__time_duration.sum')]
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        uw = units[hdr.index('dram__bytes_write.sum')]
        tm = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1}[ut]
        tot = rd * mult[ur] + wr * mult[uw]
        print(f'  => traffic {tot/1e6:.1f} MB, {tot/(t*tm)/1e9:.0f} GB/s DRAM')
    except Exception as e:
        pass
