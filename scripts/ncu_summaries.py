"""Reads ncu reports brought back in gpurun_out/ (ncu -i ... --page raw --csv) and writes the per-kernel summaries that are kept under
profiles/ + profiles/icp_reduce_dram_traffic.json (what bench.py reports as roofline.traffic).
    python scripts/ncu_summaries.py gpurun_out/r2a_prof_icp.ncu-rep gpurun_out/r2a_prof_tracker.ncu-rep"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def main(reps):
    for rep in reps:
        hdr, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        name = os.path.basename(rep).replace(".ncu-rep", "")
        lines = ["ncu --set full --clock-control none (cold caches, one kernel at a time); report %s" % os.path.basename(rep)]
        for r in rows:
            lines.append("--- " + r[idx["Kernel Name"]][:110])
            for w in WANT:
                if w in idx and r[idx[w]] != "":
                    lines.append("  %-84s %18s %s" % (w, r[idx[w]], units[idx[w]]))
        path = os.path.join(ROOT, "profiles", name.replace("r2a_", "r2_").replace("prof_", "") + "_ncu_full.txt")
        open(path, "w").write("\n".join(lines) + "\n")
        print("wrote", path)
        if "icp" in name:
            r = rows[-1]
            to_b = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tb = to_b(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_b(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            json.dump({"dram_bytes_per_launch": tb, "kernel": r[idx["Kernel Name"]][:80], "gpu_time_us_under_ncu": float(r[idx["gpu__time_duration.sum"]]),
                       "source": "ncu --set full --clock-control none, profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum of one launch, level 0 of 640x480)" % os.path.basename(path)},
                      open(os.path.join(ROOT, "profiles", "icp_reduce_dram_traffic.json"), "w"), indent=1)
            print("wrote profiles/icp_reduce_dram_traffic.json:", tb)


if __name__ == "__main__":
    main(sys.argv[1:])
