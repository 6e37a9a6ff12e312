"""Key metrics + stall reasons of every launch in an .ncu-rep:  python tools/ncu_show.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr = rr[0]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w[:62]:62s}", rr[1][i][:10], [r[i][:40] for r in rr[2:]])
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        vals = [float(r[i]) for r in rr[2:]]
        if max(vals) > 0.2:
            print(f"  stall {h[34:-23]:28s}", [f"{v:.2f}" for v in vals])
