"""Selected metrics of an `ncu --set full` report, one row per metric, one column per captured launch.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/rNN_ncu_full_X.txt"""
import csv
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr, units, data = rows[0], rows[1], rows[2:]
for name in KEEP:
    if name in hdr:
        i = hdr.index(name)
        print(f"{name} [{units[i]}] {[r[i][:34] for r in data]}")
