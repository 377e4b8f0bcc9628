"""Dev tool: headline metrics per kernel launch from an .ncu-rep (read with the ncu CLI, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'launch__grid_size',
        'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_barriers', 'launch__occupancy_limit_warps',
        'sm__maximum_warps_per_active_cycle_pct', 'launch__waves_per_multiprocessor']
want += [h for h in hdr if 'average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
seen = set()
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')][:44]
    if name in seen and '--all' not in sys.argv:
        continue
    seen.add(name)
    print('----', name)
    for w in want:
        if w not in hdr:
            continue
        i = hdr.index(w)
        try:
            if float(vals[i].replace(',', '')) < 0.2 and 'stalled' in w:
                continue
        except ValueError:
            pass
        print(f'  {w.replace("smsp__average_warps_issue_stalled_", "stall_"):66s} {vals[i]:>16s} {units[i]}')
