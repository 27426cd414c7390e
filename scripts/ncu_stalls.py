"""Print the warp-stall breakdown, memory and pipe metrics of launch N of an .ncu-rep."""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + idx]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum"]
print(data[hdr.index("Kernel Name")][:100])
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"  {k} = {data[i]} {units[i]}")
tot = 0; st = []
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        v = float(data[i] or 0); tot += v; st.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
for v, h in sorted(st, reverse=True)[:10]:
    print(f"  stall {h:24s} {v/tot*100:5.1f}%")
