#!/bin/bash
# Final single-GPU session of round 2 (after the sweep / Jacobi / relax rewrites): suite, smoke, bench line, launch list.
O=gpurun_out/final2; mkdir -p $O
python -m pytest tests -m gpu -q > $O/r2_gputest_final.log 2>&1; tail -3 $O/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 > $O/bench_r2_final.json 2> $O/bench_r2_final.err; python3 -c "
import json; d=json.loads(open('$O/bench_r2_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'fixed', d['fixed_boundary']['value'], 'pw', d['plasma_wall']['value'], 'gemm', d['plasma_wall']['roofline_gemm']['achieved'], d['plasma_wall']['roofline_gemm']['frac'], 'cpu', d['cpu_baseline']['value'], 'parity', d['parity_checked'], 'frac', d['roofline']['frac'], 'fp64', d['roofline'].get('fp64',{}).get('frac'), 'traffic', d['roofline'].get('traffic'), 'launches', d['gpu_launches'])"
tail -2 $O/bench_r2_final.err
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; tail -1 $O/ncu_launches.log | cut -c1-200
python tools/bench_configs.py > $O/r2_configs_final.log 2>&1; cat $O/r2_configs_final.log
