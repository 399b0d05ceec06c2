"""print the key metrics of every kernel in an .ncu-rep (needs only `ncu -i`, no GPU)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'local_load_bytes', 'lts__t_bytes.sum']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            print('   %-62s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h and r[i].replace('.', '').isdigit()]
    st.sort(reverse=True)
    tot = sum(s for s, _ in st) or 1
    for s_, h in st[:8]:
        print('      %5.1f%% %s' % (100 * s_ / tot, h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
