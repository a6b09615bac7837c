"""Per-kernel summary table of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
idx = {k: i for i, k in enumerate(h)}
cols = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("smsp__inst_executed.sum", "warp-inst"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2 hit%"), ("l1tex__t_sector_hit_rate.pct", "L1 hit%")]
print("%-44s %-14s" % ("kernel", "grid") + "".join("%16s" % c[1] for c in cols))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0][-44:]
    out = "%-44s %-14s" % (name, r[idx["Grid Size"]].replace(" ", ""))
    for k, _ in cols:
        v = r[idx[k]] if k in idx else ""
        u = units[idx[k]] if k in idx else ""
        try:
            f = float(v.replace(",", ""))
            v = ("%.3g" % f) if abs(f) < 1e5 else ("%.4g" % f)
        except ValueError:
            pass
        out += "%16s" % (v + (" " + u if u in ("us", "ns", "ms", "Mbyte", "Kbyte", "byte", "Gbyte") else ""))
    print(out)
