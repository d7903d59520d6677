"""Text summary of an .ncu-rep (key metrics + stall reasons per kernel): python tests/dev/ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("  Kernel Name".ljust(73), d["Kernel Name"])
        for k in KEEP:
            if k in d and d[k] != "":
                print(("  " + k).ljust(73), d[k], u.get(k, ""))
        st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
              for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
        print("  stall reasons (warps per issue-active cycle)".ljust(73), ", ".join("%s %.2f" % kv for kv in top))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
