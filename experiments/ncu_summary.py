"""Print the judged metrics of every kernel instance in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'launch__waves_per_multiprocessor']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("%s  grid %s block %s" % (r[col['Kernel Name']][:100], r[col['Grid Size']], r[col['Block Size']]))
    for w in WANT:
        if w in col:
            print("    %-68s %14s %s" % (w, r[col[w]], units[col[w]]))
