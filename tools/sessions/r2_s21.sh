#!/bin/bash
# session 21: register-carried sweep variant (GSB_SWEEP_UNROLL=1): parity
O=gpurun_out/s21; mkdir -p $O
GSB_SWEEP_UNROLL=1 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_slab.py tests/test_gpu_streaming_parity.py -q -m gpu > $O/tests.log 2>&1
tail -12 $O/tests.log
