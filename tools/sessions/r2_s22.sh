#!/bin/bash
# session 22: register-carried sweep kernel as the default: full suite, smoother bench, caller benches, ncu capture
O=gpurun_out/s22; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
FUSE_ONLY=3 python tools/bench_smooth.py 129x129x4096 257x257x256 513x513x256 1025x1025x64 4097x4097x1 8193x8193x1 2>&1 | grep fuse | tee $O/smooth.log
python tools/bench_batch257.py 257 256 2>&1 | tail -1
python tools/bench_batch257.py 513 128 2>&1 | tail -1
python tools/bench_configs.py > $O/configs.log 2>&1; cat $O/configs.log
python bench.py --workload slab --steps 5 2>/dev/null > $O/slab4097.json; python -c "
import json; d=json.loads(open('$O/slab4097.json').read().strip().splitlines()[-1]); print('slab 4097', d['value'], 'GLUPS', d['ms_per_vcycle'], 'ms/vcycle frac', d['roofline']['frac'])"
FUSE_ONLY=3 ncu --set full --clock-control none --import-source on -k regex:k_sweep_warp -s 2 -c 1 -f -o $O/sweep_rc python tools/bench_smooth.py 8193x8193x1 > $O/ncu.log 2>&1; tail -1 $O/ncu.log
