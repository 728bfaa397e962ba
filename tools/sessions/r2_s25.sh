#!/bin/bash
# session 25: warp-per-30-columns k_relax: parity of every streaming-path test, then the batches
O=gpurun_out/s25; mkdir -p $O
python -m pytest tests/test_gpu_streaming_parity.py tests/test_gpu_picard.py tests/test_gpu_anderson.py tests/test_gpu_free_boundary_batched.py tests/test_gpu_free_boundary.py -q -m gpu -x > $O/tests.log 2>&1; tail -3 $O/tests.log
python tools/bench_batch257.py 257 256 2>&1 | tail -1
python tools/bench_batch257.py 513 128 2>&1 | tail -1
python tools/bench_configs.py 2>&1 | tail -6
