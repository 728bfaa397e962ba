"""Stall-reason breakdown (ncu warp-state samples) per outermost source-line range.
usage: python tools/ncu_stalls.py report.ncu-rep object.o kernel name,file,lo,hi ..."""
import csv, subprocess, sys, re, collections, io, os, tempfile
rep, obj, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}; cur = None; infn = False
for ln in dis.splitlines():
    if re.match(r"\s*\.section\s+\.text\.", ln) or ln.startswith(".text."):
        infn = kern in ln; cur = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia = hdr.index("Address")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
specs = []
for spec in sys.argv[4:]:
    name, f, a, b = spec.split(","); specs.append((name, f, int(a), int(b), collections.Counter()))
for r in rows[2:]:
    try: addr = int(r[ia], 16)
    except (ValueError, IndexError): continue
    if base is None: base = addr
    key = line_of.get(addr - base)
    if not key: continue
    for name, f, a, b, cnt in specs:
        if key[0] == f and a <= key[1] <= b:
            for i, h in stall_cols:
                cnt[h] += int(r[i] or 0)
for name, f, a, b, cnt in specs:
    tot = sum(cnt.values()) or 1
    print(f"{name:12s} total {tot:8d}: " + "  ".join(f"{h[6:]}={100*v/tot:.0f}%" for h, v in cnt.most_common(7)))
