"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-step table."""
import collections, csv, sys

def main(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"]) / 1000.0, r.get("Grid Size", "")) for r in rows]
    g = [i for i, (n, _, _) in enumerate(names) if "k_gather_rows" in n]
    # one resident step = from one pair of gathers to the next pair
    starts = [i for k, i in enumerate(g) if k == 0 or g[k - 1] != i - 1]      # first gather of each pair
    s, e = starts[-2], starts[-1]
    step = names[s:e]
    tot = sum(t for _, t, _ in step)
    agg = collections.OrderedDict()
    for n, t, _ in step:
        k = n.split("(")[0].replace("void ", "").replace("ader::", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
    with open(out, "w") as f:
        f.write("# Launch list of one bench step (ncu gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Source: `%s`. Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n" % path)
        f.write("Launches in the step: %d, sum %.1f us\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n" % (len(step), tot))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, c, t, 100 * t / tot))
    print(open(out).read())

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
