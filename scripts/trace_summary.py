"""Summarise a chrome trace exported by bench.py (ADER_B200_TRACE=...): per-step kernel timeline of the LAST profiled
step (start offset, duration, stream) so the critical path of the fork/join DAG can be read off."""
import json, sys
ev = json.load(open(sys.argv[1]))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
# one step = from one k_gather_batch to the next
starts = [i for i, e in enumerate(ks) if "k_gather_batch" in e["name"]]
if len(starts) >= 2:
    a, b = starts[-2], starts[-1]
else:
    a, b = 0, len(ks)
step = ks[a:b]
t0 = step[0]["ts"]
print("step span %.1f us, %d kernels, busy sum %.1f us" % (step[-1]["ts"] + step[-1]["dur"] - t0, len(step), sum(e["dur"] for e in step)))
for e in step:
    n = e["name"].split("(")[0].replace("void ", "").replace("ader::", "")
    print("%8.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), n[:60]))
