#!/bin/bash
# session 20: ncu --set full of the new k_sweep_warp<6,2> on 8193^2 (same shape as the r2 capture sweep_warp6_b)
O=gpurun_out/s20; mkdir -p $O
FUSE_ONLY=3 python tools/bench_smooth.py 8193x8193x1 > $O/plain.log 2>&1 && \
FUSE_ONLY=3 ncu --set full --clock-control none --import-source on -k regex:k_sweep_warp -s 2 -c 1 -f -o $O/sweep_rot python tools/bench_smooth.py 8193x8193x1 > $O/ncu.log 2>&1
tail -2 $O/ncu.log; cat $O/plain.log
