"""Print, for each kernel in an .ncu-rep, duration / DRAM bytes / top stall-sampled SASS lines (needs ncu on PATH)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
seen = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    if pat and pat not in name:
        continue
    if name in seen:
        continue
    seen[name] = r[ix["ID"]]
    g = lambda k: r[ix[k]] if k in ix else "?"
    print("== %s id=%s grid=%s dur=%sus dram_rd=%sMB dram_wr=%sMB regs=%s occ=%s%% issue_active=%s%% l2hit=%s%%" % (
        name, r[ix["ID"]], g("launch__grid_size"), g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
        g("launch__registers_per_thread"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("lts__t_sector_hit_rate.pct")))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:1" % name.split("::")[-1].split("<")[0]],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 3:
        continue
    sh = srows[1]; six = {n: i for i, n in enumerate(sh)}
    data = []
    for q in srows[2:]:
        try:
            data.append((int(q[six["# Samples"]] or 0), q[six["Source"]][:100]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    for smp, s in sorted(data, reverse=True)[:8]:
        print("     %5.1f%%  %s" % (100.0 * smp / tot, s))
