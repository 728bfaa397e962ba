#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_picard.py tests/test_gpu_streaming_parity.py tests/test_gpu_free_boundary_batched.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_batch257.py 257 256
python tools/bench_batch257.py 513 128
