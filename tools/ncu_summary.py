"""Prints the key metrics of every kernel in an .ncu-rep (needs ncu on PATH; no GPU)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        # issue utilisation per execution pipe (ALU = integer / DPX min-max pipe, FMA = IMAD pipe, LSU = shared / global / shuffle)
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=====", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for w in WANT:
            if w in hdr:
                print("  %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = [(float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(reverse=True)
        print("  stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
