"""Summarise an .ncu-rep into a small text file for profiles/ (run in the build container)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"# {rep}\n# kernel: {name}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"{w:72s} {vals[i]:>20s} {units[i]}\n")
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                    st.append((float(vals[i].replace(",", "")),
                               h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            f.write("stall cycles per issued instruction: " + ", ".join(f"{h}={v:.2f}" for v, h in sorted(st, reverse=True)[:9]) + "\n\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
