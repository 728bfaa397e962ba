#!/bin/bash
# session 19: effect of the small-body sweep kernel on the callers (full suite, streaming Picard batches, slab V-cycle)
O=gpurun_out/s19; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
python tools/bench_batch257.py 257 256 2>&1 | tail -2
python tools/bench_batch257.py 513 128 2>&1 | tail -2
python tools/bench_configs.py > $O/configs.log 2>&1; cat $O/configs.log
python bench.py --workload slab --steps 5 2>/dev/null > $O/slab4097.json; cut -c1-200 $O/slab4097.json
