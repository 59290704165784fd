"""Summarises an .ncu-rep (read here, no GPU needed): per kernel launch the metrics the roofline discussion uses.
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep ["command line that produced it"] > profiles/rNN_....txt"""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2:
    print(sys.argv[2])
print("(cold-cache, serialised launches under the profiler: shares and byte counts are meaningful, absolute times are "
      "not bench numbers)\n")
for r in data:
    print(r[col["Kernel Name"]])
    for m in METRICS:
        if m in col and r[col[m]] != "":
            print("    %-70s %s %s" % (m, r[col[m]], units[col[m]]))
    print()
