"""Prints the handful of counters that decide what bounds a kernel from an `ncu --page raw --csv` export:
python tools/ncu_summary.py gpurun_out/x_full_raw.csv [...]"""
import csv, sys
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
STALL = 'smsp__average_warps_issue_stalled_'
for fn in sys.argv[1:]:
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(fn)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i][:110]} {units[i]}")
        st = sorted(((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                     if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')), reverse=True)
        print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
