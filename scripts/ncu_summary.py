"""Summarise an `ncu --page raw --csv` export: one line per launch with the counters the roofline discussion uses.
Usage: python scripts/ncu_summary.py raw.csv > summary.md"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def get(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    except ValueError:
        return default


print("| kernel | grid | regs | us | DRAM rd MB | DRAM wr MB | L2->SM MB | tensor pipe % (elapsed) | tcgen05 (utc) pipe % | legacy hmma % | warps active % |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("ader::", "")
    tc = get(r, "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active")
    utc = max(get(r, n, 0.0) for n in hdr if "pipe_tensor_subpipe_utc" in n and n.endswith("pct_of_peak_sustained_elapsed")) if any("subpipe_utc" in n for n in hdr) else float("nan")
    print("| `%s` | %d | %d | %.1f | %.2f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f |" % (
        name[:48], get(r, "launch__grid_size", 0), get(r, "launch__registers_per_thread", 0), get(r, "gpu__time_duration.sum"),
        get(r, "dram__bytes_read.sum") / 1e6, get(r, "dram__bytes_write.sum") / 1e6, get(r, "l1tex__m_xbar2l1tex_read_bytes.sum") / 1e6,
        get(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"), utc,
        get(r, "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"), get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
