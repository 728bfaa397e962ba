#!/bin/bash
# session 26: where k_relax stands now (launch list + one full capture)
O=gpurun_out/s26; mkdir -p $O
GSB_NO_GRAPH=1 python tools/bench_batch257.py 257 256 12 > $O/plain.log 2>&1 && \
GSB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches257.csv python tools/bench_batch257.py 257 256 12 > $O/ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/s26/launches257.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
    print(f"{k[:60]:60s} n={n:4d} total {v/1e3:9.1f} us  avg {v/n/1e3:8.1f} us  {100*v/tot:5.1f} %")
PY
GSB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_relax|k_topo|k_source_raw" -s 9 -c 3 -f -o $O/stencils2 python tools/bench_batch257.py 257 256 12 > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log
