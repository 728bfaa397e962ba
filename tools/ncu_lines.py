"""Attribute ncu per-instruction samples of one kernel to CUDA source lines.
usage: python tools/ncu_lines.py report.ncu-rep object.o mangled_kernel_name [top]
Needs the object built with -lineinfo; uses `nvdisasm -g` for the SASS-offset -> line map."""
import csv, subprocess, sys, re, collections, io, os, tempfile
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
# parse: within the kernel's .text section, lines "//## File "...", line N" then instructions "/*0010*/ OP ..."
line_of = {}
cur = None; infn = False; inl = []
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = kern in ln
        cur = None
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:  # with -gi an inline chain is printed innermost first; the last marker is the outermost frame
        cur = (os.path.basename(m.group(1)), int(m.group(2)), False)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0])
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= max(ie, isamp): continue
    try: addr = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = addr
    off = addr - base
    key = line_of.get(off, ("?", 0, False))
    s = int(r[isamp] or 0); e = int(r[ie] or 0)
    agg[key[:2]][0] += s; agg[key[:2]][1] += e
    tot_s += s; tot_e += e
print(f"total samples {tot_s}, warp instructions {tot_e}")
for (f, l), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:5d}  samples {s:7d} ({100*s/tot_s:5.1f}%)  inst {e:10d} ({100*e/tot_e:5.1f}%)")
