"""Key metrics of every launch in an `ncu --page raw --csv` export."""
import csv, sys
rd = list(csv.reader(open(sys.argv[1])))
hdr, units = rd[0], rd[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k[:80]:80s} {units[i]:10s}", [r[i][:40] for r in rd[2:]])
