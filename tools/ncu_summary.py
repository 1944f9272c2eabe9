"""Key counters of an `ncu --set full` report, one column per captured launch.  usage: python tools/ncu_summary.py x.ncu-rep"""
import csv, subprocess, sys
WANT = ["Grid Size", "Block Size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("kernel:", r[col["Kernel Name"]][:110])
print("%-66s %-10s %s" % ("metric", "unit", "  ".join("launch %d" % i for i in range(len(data)))))
for w in WANT:
    if w in col:
        print("%-66s %-10s %s" % (w, units[col[w]], "  ".join(r[col[w]] for r in data)))
