#!/bin/bash
# session 17: Anderson mixing on the device - parity tests, then the whole GPU suite
mkdir -p gpurun_out/s17
python -m pytest tests/test_gpu_anderson.py -q -m gpu > gpurun_out/s17/anderson.log 2>&1
tail -15 gpurun_out/s17/anderson.log
python -m pytest tests -q -m gpu -x > gpurun_out/s17/tests.log 2>&1
tail -5 gpurun_out/s17/tests.log
