#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_batch257.py 257 256
python tools/bench_batch257.py 513 128
python bench.py --workload slab --steps 5 2>/dev/null | cut -c1-160
python bench.py --workload slab --grid 8193 --steps 3 2>/dev/null | cut -c1-160
python tools/bench_smooth.py 2>&1 | tail -8
