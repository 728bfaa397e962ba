#!/bin/bash
# session 28: run-to-run spread of the one-GPU slab line (0.957 vs 1.100 ms per V-cycle in two sessions on the same code path)
O=gpurun_out/s28; mkdir -p $O
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm,clocks.mem --format=csv
for i in 1 2 3; do
python bench.py --workload slab --steps 10 2>/dev/null > $O/slab4097_$i.json; python -c "
import json; d=json.loads(open('$O/slab4097_$i.json').read().strip().splitlines()[-1]); print('slab 4097 run $i', round(d['value'],1), 'GLUPS', round(d['ms_per_vcycle'],4), 'ms/vcycle frac', round(d['roofline']['frac'],3))"
done
FUSE_ONLY=3 python tools/bench_smooth.py 4097x4097x1 2049x2049x1 1025x1025x1 513x513x1 257x257x1 2>&1 | grep fuse
python bench.py --workload slab --grid 8193 --steps 5 2>/dev/null > $O/slab8193.json; python -c "
import json; d=json.loads(open('$O/slab8193.json').read().strip().splitlines()[-1]); print('slab 8193', round(d['value'],1), 'GLUPS', round(d['ms_per_vcycle'],4), 'ms/vcycle frac', round(d['roofline']['frac'],3))"
