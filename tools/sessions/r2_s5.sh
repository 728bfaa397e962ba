#!/bin/bash
O=gpurun_out
show() { python3 -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', 'ms/solve', round(d['ms_per_step'],2), 'ms/vcycle', round(d['ms_per_vcycle'],3), 'launches', d['gpu_launches'])"; }
python bench.py --workload slab --steps 5 > $O/d1.json 2>/dev/null; show $O/d1.json direct
OMP_NUM_THREADS=1 python bench.py --workload slab --steps 5 > $O/d2.json 2>/dev/null; show $O/d2.json omp1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --workload slab --steps 5 > $O/d3.json 2>/dev/null; show $O/d3.json torchrun
python bench.py --workload slab --steps 5 > $O/d4.json 2>/dev/null; show $O/d4.json direct_again
python -m pytest tests/test_gpu_picard.py tests/test_gpu_multigrid.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py > $O/plain_cfgs.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_cfgs_launches.csv python tools/bench_configs.py > $O/ncu_cfgs.log 2>&1; tail -3 $O/ncu_cfgs.log
