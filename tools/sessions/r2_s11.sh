#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/latency_probe.py
python bench.py --steps 3 > $O/r2_bench11.json 2> $O/r2_bench11.err; python3 -c "
import json; d=json.loads(open('$O/r2_bench11.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'fixed', d['fixed_boundary']['value'], 'pw', d['plasma_wall']['value'], 'gemm', d['plasma_wall']['roofline_gemm']['achieved'], d['plasma_wall']['roofline_gemm']['frac'], 'cpu', d['cpu_baseline']['value'], 'parity', d['parity_checked'], 'fp64', d['roofline'].get('fp64',{}).get('frac'))"
tail -2 $O/r2_bench11.err
