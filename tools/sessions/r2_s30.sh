#!/bin/bash
# session 30: flat-index loops in k_topo / k_source_raw / k_source_scale (no single-thread last trip at nr = 2^k + 1)
O=gpurun_out/s30; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
python tools/bench_batch257.py 257 256 2>&1 | tail -1
python tools/bench_batch257.py 513 128 2>&1 | tail -1
python tools/bench_configs.py 2>&1 | tail -4
