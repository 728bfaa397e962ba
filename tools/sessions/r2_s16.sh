#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --workload fixed_boundary --steps 3 --no-extras --no-cpu-baseline 2>/dev/null | python3 -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fixed-boundary eq/s', round(d['value']), 'converged', d['stats']['converged'])"; done
python bench.py --workload fixed_boundary --steps 1 --no-extras --no-cpu-baseline > $O/plain_fixed.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_picard_resident -s 3 -c 1 python bench.py --workload fixed_boundary --steps 1 --no-extras --no-cpu-baseline 2>&1 | grep -E "dram__|gpu__time|lts__|fp64" 
