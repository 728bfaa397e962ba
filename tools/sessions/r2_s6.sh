#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_configs.py 2>&1 | tail -6
GSB_NO_GRAPH=1 python tools/bench_configs.py 2>&1 | tail -6 | head -3
python tools/bench_batch257.py 257 256 16 > $O/plain_b257.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_b257_launches.csv python tools/bench_batch257.py 257 256 16 > $O/ncu_b257.log 2>&1; tail -2 $O/ncu_b257.log
