"""profiles/ summaries of the round-2 staged pipeline from the reports scripts/gpu_profile2.sh brings back:
    python scripts/ncu_frame_summary.py gpurun_out/r2c_launches.csv gpurun_out/r2c_prof_frame.ncu-rep r2"""
import csv, io, os, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]

# ---- launch list: one steady-state frame = the launches between two consecutive tracker launches
rows = [r for r in csv.reader(open(launch_csv, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
launches = [(r[iK].split("(")[0].replace("void ", "").strip(), float(r[iV].replace(",", "")) / 1e3) for r in rows if r is not hdr and r[iM] == "gpu__time_duration.sum"]
tr = [i for i, (k, _) in enumerate(launches) if "track_persistent" in k]
a, b = tr[-3], tr[-2]
frame = launches[a:b]
agg = OrderedDict()
for k, t in frame:
    n, s = agg.get(k, (0, 0.0)); agg[k] = (n + 1, s + t)
tot = sum(t for _, t in frame)
out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 6 --warmup 3 --sequences 1 --extras 0 (round 2, staged pipeline)",
       "one steady-state frame (the launches between two tracker launches, ~300 k surfels): %d launches, %.1f us serialised (cold caches, one kernel at a time: compare shares)" % (len(frame), tot)]
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("  %-34s x%-2d %8.1f us  %4.1f%%" % (k[:34], n, s, 100 * s / tot))
staged = ("depth_filter_metric", "vertex_normal_radius", "curvature_gradient", "sobel_cand", "fuse_normals", "so3_image", "so3_prealign")
st = sum(s for k, (n, s) in agg.items() if any(x in k for x in staged))
out.append("  of which run one frame ahead on the staging streams (camera frame only): %.1f us + the current-frame half of prep_all; so3_prealign_kernel is ONE CTA (latency, not work)" % st)
p = os.path.join(ROOT, "profiles", tag + "_launches_summary_single_sequence_staged.txt")
open(p, "w").write("\n".join(out) + "\n"); print("wrote", p)

# ---- per-kernel table from the full report
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
h, rows = r[0], r[2:]
idx = {x: i for i, x in enumerate(h)}
cols = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "warp-inst"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st.long"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st.math"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st.notsel"), ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st.bar")]
out = ["ncu --set full --clock-control none (cold caches, one kernel at a time), every kernel of one steady-state frame except the tracker; report %s" % os.path.basename(rep),
       "%-30s" % "kernel" + "".join("%11s" % c[1] for c in cols)]
seen = set()
for row in rows:
    name = row[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    key = (name, row[idx["launch__grid_size"]] if "launch__grid_size" in idx else "")
    def fmt(c):
        v = row[idx[c]] if c in idx else ""
        try: return "%11.4g" % float(v.replace(",", ""))
        except ValueError: return "%11s" % "-"
    out.append("%-30s" % name[:29] + "".join(fmt(c[0]) for c in cols))
p = os.path.join(ROOT, "profiles", tag + "_frame_kernels_ncu_full.txt")
open(p, "w").write("\n".join(out) + "\n"); print("wrote", p)
