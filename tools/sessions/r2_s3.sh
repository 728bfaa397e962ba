#!/bin/bash
# GPU session 3 (1 x B200): regression, bench, slab launch list, GEMM capture
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_gputest3.log 2>&1; tail -4 $O/r2_gputest3.log
python bench.py --steps 3 --warmup 3 > $O/r2_bench3.json 2> $O/r2_bench3.err; cut -c1-400 $O/r2_bench3.json; tail -3 $O/r2_bench3.err
python bench.py --workload slab --steps 3 > $O/r2_slab1.json 2> $O/r2_slab1.err; cat $O/r2_slab1.json | cut -c1-900; tail -3 $O/r2_slab1.err
python bench.py --workload slab --steps 1 > $O/plain_slab.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_slab_launches.csv python bench.py --workload slab --steps 1 > $O/ncu_slab.log 2>&1; tail -2 $O/ncu_slab.log
python tools/bench_wall_gemm.py 129 4096 > $O/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wall_gemm_sk -s 2 -c 1 -f -o $O/r2_wall_gemm_v8 python tools/bench_wall_gemm.py 129 4096 > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log
