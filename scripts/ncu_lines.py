"""Per-source-line profile of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py REP KERNEL_REGEX [top]  -> lines ranked by stall samples, with instructions executed."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                      "::regex:%s:1" % pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or r[2] != "-":          # keep the CUDA-line aggregate rows only (Address == "-")
        continue
    try:
        data.append((int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0), cur_file, r[0], r[1].strip()[:110]))
    except (ValueError, KeyError):
        pass
ts = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
for s, n, f, ln, src in sorted(data, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100.0 * s / ts, 100.0 * n / ti, f, ln, src))
