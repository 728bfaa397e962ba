#!/bin/bash
# session 23: temporally blocked Jacobi seed: parity, then the streaming Picard batches
O=gpurun_out/s23; mkdir -p $O
python -m pytest tests/test_gpu_jacobi_blocked.py -q -m gpu > $O/jac.log 2>&1; tail -15 $O/jac.log
python -m pytest tests/test_gpu_streaming_parity.py tests/test_gpu_picard.py -q -m gpu -x > $O/tests.log 2>&1; tail -3 $O/tests.log
python tools/bench_batch257.py 257 256 2>&1 | tail -1
GSB_NO_FUSED_JACOBI=1 python tools/bench_batch257.py 257 256 2>&1 | tail -1
python tools/bench_batch257.py 513 128 2>&1 | tail -1
