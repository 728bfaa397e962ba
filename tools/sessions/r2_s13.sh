#!/bin/bash
O=gpurun_out
for T in 256 384; do
  if [ $T = 512 ]; then L=$PWD/scpn_fusion_core_b200/libgsb200.so; else L=$PWD/scpn_fusion_core_b200/libgsb200_t$T.so; fi
  echo "== threads $T"
  GSB200_LIB=$L python -m pytest tests/test_gpu_picard.py tests/test_gpu_multigrid.py -m gpu -x -q 2>&1 | tail -1
  GSB200_LIB=$L python bench.py --workload fixed_boundary --steps 3 --no-extras --no-cpu-baseline 2>/dev/null | python3 -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fixed-boundary eq/s', round(d['value']), 'converged', d['stats']['converged'])"
done
