"""Where a kernel's warp-stall samples and executed instructions sit, by SASS region (ncu --set full --import-source on report):
    python scripts/ncu_sass_regions.py report.ncu-rep kernel_regex [region_size]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# several launches may match: take the first block
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
hdr, data = rows[start], rows[start + 1:end]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
num = lambda v: int(v) if v.isdigit() else 0
tot, toti = sum(num(r[iS]) for r in data), sum(num(r[iI]) for r in data)
print(rows[start - 1][1][:80], "| samples", tot, "| warp instr %.2fM" % (toti / 1e6), "| SASS lines", len(data))
for a in range(0, len(data), step):
    blk = data[a:a + step]
    s, ins = sum(num(r[iS]) for r in blk), sum(num(r[iI]) for r in blk)
    if ins == 0: continue
    ops = {}
    for r in blk:
        t = r[iSrc].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")).split(".")[0]
        ops[op] = ops.get(op, 0) + num(r[iI])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(f"  sass {a:4d}-{a + step - 1:4d}: samples {100 * s / tot:5.1f}%  instr {100 * ins / toti:5.1f}%  " + " ".join(f"{k}:{v / 1e6:.2f}M" for k, v in top))
