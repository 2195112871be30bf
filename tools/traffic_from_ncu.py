"""ncu --set full capture of ONE frame (tools/profile_frame.py --iters 1 under `ncu --set full -k regex:k_ ...`) -> profiles/traffic.json.

    python tools/traffic_from_ncu.py <capture.ncu-rep> <config key, e.g. c1_k20> [--last N] [--out file.json]

For every launch of the LAST frame in the capture (N launches; default: all), in launch order: the kernel symbol, measured DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum), duration under ncu, and the utilisation of the units that can bound it.  `bound` is the
busiest of them: hbm (DRAM throughput), l1tex (L1 / shared-memory data path: LSU wavefronts, REDs, LDGSTS), alu (the integer / DPX pipe),
issue (warp schedulers).  bench.py pairs these records with its own event-timed launches of the same frame, in order, and refuses to print a
`traffic` for a configuration without a capture."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {
    "time_ns": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "hbm": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}


def records(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(v) for k, v in M.items() if v in hdr}
    recs = []
    for r in rows[2:]:
        def val(k):
            if k not in col or r[col[k]] == "":
                return None
            return float(r[col[k]].replace(",", "")) * SCALE.get(units[col[k]], 1)
        name = re.sub(r"^void\s+", "", r[hdr.index("Kernel Name")])
        name = re.sub(r"\(.*\)$", "", name).strip()
        rec = {"kernel": name, "dram_bytes": int((val("rd") or 0) + (val("wr") or 0)), "ncu_us": round((val("time_ns") or 0) / 1e3, 1)}
        for k in ("hbm", "l1tex", "lts", "alu", "lsu", "issue", "warps", "regs"):
            v = val(k)
            rec[k + ("_pct" if k != "regs" else "")] = None if v is None else round(v, 1)
        cand = {k: rec[k + "_pct"] for k in ("hbm", "l1tex", "alu", "issue") if rec.get(k + "_pct") is not None}
        rec["bound"] = max(cand, key=cand.get) if cand else None
        recs.append(rec)
    return recs


def main():
    path, key = sys.argv[1], sys.argv[2]
    recs = records(path)
    if "--last" in sys.argv:
        recs = recs[-int(sys.argv[sys.argv.index("--last") + 1]):]
    tj = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tj) as f:
            data = json.load(f)
    except Exception:
        data = {}
    data["_comment"] = ("per launch of one frame, in launch order, from `ncu --set full --clock-control none` captures (tools/traffic_from_ncu.py): measured DRAM bytes "
                        "and the utilisation of the units that can bound the kernel; bench.py pairs them with its event-timed launches")
    data[key] = {"capture": os.path.basename(path), "launches": recs}
    with open(tj, "w") as f:
        json.dump(data, f, indent=1)
    for r in recs:
        print("%-48s %8.1f us  dram %7.3f GB  hbm %5s l1tex %5s alu %5s lsu %5s issue %5s warps %5s regs %s -> %s"
              % (r["kernel"][:48], r["ncu_us"], r["dram_bytes"] / 1e9, r["hbm_pct"], r["l1tex_pct"], r["alu_pct"], r["lsu_pct"], r["issue_pct"], r["warps_pct"], r["regs"], r["bound"]))


if __name__ == "__main__":
    main()
