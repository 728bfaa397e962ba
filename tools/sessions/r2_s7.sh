#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_batch257.py 257 256
python tools/bench_batch257.py 513 128
python tools/bench_configs.py 2>&1 | tail -2
python bench.py --workload slab --steps 5 2>/dev/null | cut -c1-160
python tools/bench_batch257.py 257 256 16 > $O/plain_b257.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_b257_launches2.csv python tools/bench_batch257.py 257 256 16 > $O/ncu_b257.log 2>&1; tail -1 $O/ncu_b257.log
