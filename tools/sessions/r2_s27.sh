#!/bin/bash
# session 27: more loads in flight in k_source_scale / k_copy_from_cur / k_prolong_add, more k_relax bands: suite + batches
O=gpurun_out/s27; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
python tools/bench_batch257.py 257 256 2>&1 | tail -1
python tools/bench_batch257.py 513 128 2>&1 | tail -1
python tools/bench_configs.py 2>&1 | tail -3
python bench.py --workload slab --steps 5 2>/dev/null > $O/slab4097.json; python -c "
import json; d=json.loads(open('$O/slab4097.json').read().strip().splitlines()[-1]); print('slab 4097', d['value'], 'GLUPS', d['ms_per_vcycle'], 'ms/vcycle frac', d['roofline']['frac'])"
