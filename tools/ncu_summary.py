"""Text summary of an `ncu --set full` report for profiles/: python tools/ncu_summary.py report.ncu-rep [launch index]
Prints the launch configuration, the speed-of-light / pipe / memory metrics the DESIGN quotes, the shared-memory data-pipe
wavefronts (tensor-core operand reads vs LSU) and the DRAM bytes, straight from `ncu -i ... --page raw --csv`."""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr, units, data = rows[0], rows[1], rows[2:]
d = data[idx]
col = {h: i for i, h in enumerate(hdr)}


def get(name):
    i = col.get(name)
    return None if i is None else (d[i], units[i])


print("report: %s   launch %d of %d" % (rep.split("/")[-1], idx if idx >= 0 else len(data) + idx, len(data)))
print("kernel: %s" % d[col["Kernel Name"]][:150])
for k in ("Grid Size", "Block Size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
          "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"):
    if k in col:
        print("  %-62s %s %s" % (k, d[col[k]], units[col[k]]))
WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
for k in WANT:
    v = get(k)
    if v is not None and v[0] != "":
        print("  %-82s %s %s" % (k, v[0], v[1]))
print("  warp issue stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            x = float(d[col[h]])
        except ValueError:
            continue
        if x >= 0.3:
            print("    %-40s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], x))
