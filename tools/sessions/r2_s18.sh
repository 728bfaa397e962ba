#!/bin/bash
# session 18: k_sweep_warp with rotating slot registers (small loop body) - parity + A/B of the unroll factor
mkdir -p gpurun_out/s18
python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_slab.py -q -m gpu -x > gpurun_out/s18/tests.log 2>&1
tail -3 gpurun_out/s18/tests.log
for u in 4 2; do
  echo "== GSB_SWEEP_UNROLL=$u"
  GSB_SWEEP_UNROLL=$u FUSE_ONLY=3 python tools/bench_smooth.py 257x257x256 513x513x256 1025x1025x64 4097x4097x1 8193x8193x1 2>&1 | grep fuse
done | tee gpurun_out/s18/unroll_rot.log
