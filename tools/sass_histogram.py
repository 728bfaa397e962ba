"""SASS opcode histogram per kernel of libgsb200.so (cuobjdump -sass), written as a markdown table.
usage: python tools/sass_histogram.py [lib.so] > profiles/rN_sass_histograms.md
Evidence for: which tensor / async-copy / barrier instructions each kernel really contains (DMMA, LDGSTS,
SYNCS = mbarrier, UTMALDG = TMA, BAR), FP64 instruction counts, and that only sm_100a code objects are shipped."""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "scpn_fusion_core_b200", "libgsb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern = None
hist = collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
KEY = ["DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "SYNCS", "UTMALDG", "UTMASTG", "BAR", "LDS", "STS", "LDG", "STG", "SHFL",
       "MUFU"]
print(f"# SASS opcode histograms of `{os.path.basename(lib)}` (static instruction counts; code objects: {', '.join(archs)})\n")
print("| kernel | total | " + " | ".join(KEY) + " |")
print("|---|---:|" + "---:|" * len(KEY))
for (k, h), name in zip(hist.items(), demangle):
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("gsb::", "")
    short = re.sub(r"GemmCfg<([^>]*)>", r"GemmCfg<\1>", short)
    print(f"| `{short[:70]}` | {sum(h.values())} | " + " | ".join(str(h.get(x, 0)) for x in KEY) + " |")
print("\nDMMA = mma.sync.m8n8k4.f64 (FP64 tensor pipe); LDGSTS = cp.async; SYNCS = mbarrier operations; UTMALDG/UTMASTG = TMA "
      "bulk tensor copies (none: every operand row of this path is only 8-byte aligned, see DESIGN.md); BAR = CTA barriers.")
