"""SASS opcode summary of the built library (proof of tcgen05 / TMEM / TMA use): writes profiles/r2/sass_opcodes.md.
Usage: python scripts/sass_summary.py [out.md]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ader_b200", "lib", "libader_b200.so")
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2", "sass_opcodes.md")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, per, tot = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); per[fn] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        per[fn][m.group(1)] += 1; tot[m.group(1)] += 1
mean = {"UTCHMMA": "tcgen05.mma (kind::f16, bf16 operands)", "UTCBAR": "tcgen05.commit -> mbarrier", "LDTM": "tcgen05.ld (TMEM -> registers)",
        "UTMALDG": "cp.async.bulk.tensor (TMA tensor-map load)", "UTMASTG": "TMA tensor-map store", "UBLKCP": "cp.async.bulk (linear bulk copy)",
        "HMMA": "legacy warp-level mma.sync (fp16 / tf32)", "SYNCS": "mbarrier ops", "ACQBULK": "griddepcontrol.wait (PDL)",
        "PREEXIT": "griddepcontrol.launch_dependents (PDL)", "UTCATOMSWS": "TMEM alloc / dealloc", "MEMBAR": "fences", "LDGSTS": "cp.async (Ampere-style)"}
demangle = lambda n: re.sub(r"\(.*", "", subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("(int)", "").replace("(ader::ev::EvMode)", "")) or n
L = ["# SASS opcode summary of `libader_b200.so` (sm_100a)", "", "`cuobjdump -sass ader_b200/lib/libader_b200.so`, opcode counts.", "", "## Whole library", "",
     "| opcode | count | meaning |", "|---|---:|---|"]
L += ["| `%s` | %d | %s |" % (op, tot[op], mean[op]) for op in mean if tot[op]]
L += ["", "## Per kernel (kernels that use tcgen05 / TMA / bulk copies / mma.sync)", "", "| kernel | UTCHMMA | LDTM | UTMALDG | UBLKCP | HMMA | total instr |",
      "|---|---:|---:|---:|---:|---:|---:|"]
for fn, c in per.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["HMMA"]:
        L.append("| `%s` | %d | %d | %d | %d | %d | %d |" % (demangle(fn)[:70], c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["HMMA"], sum(c.values())))
L += ["", "Script: `python scripts/sass_summary.py`."]
os.makedirs(os.path.dirname(dst), exist_ok=True)
open(dst, "w").write("\n".join(L) + "\n")
print("wrote", dst)
