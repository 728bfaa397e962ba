#!/bin/bash
# GPU session 2 (1 x B200): wall GEMM bench + correctness, FMA-variant parity/perf, ncu captures.
O=gpurun_out
python tools/bench_wall_gemm.py > $O/r2_gemm.json 2> $O/r2_gemm.err; tail -3 $O/r2_gemm.json | cut -c1-400; tail -3 $O/r2_gemm.err
python -m pytest tests -m gpu -q -x -k "plasma_wall or abi or free_boundary" > $O/r2_s2_tests.log 2>&1; tail -3 $O/r2_s2_tests.log
# FMA measurement build: which parity tests survive, and what it buys
GSB200_LIB=$PWD/scpn_fusion_core_b200/libgsb200_fma.so python -m pytest tests -m gpu -q > $O/r2_fma_tests.log 2>&1; tail -40 $O/r2_fma_tests.log | cut -c1-200
python bench.py --workload fixed_boundary --steps 3 --no-extras --no-cpu-baseline > $O/r2_fixed_exact.json 2>$O/r2_fixed_exact.err; cut -c1-300 $O/r2_fixed_exact.json
GSB200_LIB=$PWD/scpn_fusion_core_b200/libgsb200_fma.so python bench.py --workload fixed_boundary --steps 3 --no-extras --no-cpu-baseline > $O/r2_fixed_fma.json 2>$O/r2_fixed_fma.err; cut -c1-300 $O/r2_fixed_fma.json; tail -2 $O/r2_fixed_fma.err
# ncu: GEMM full capture, Picard kernel full capture at the headline batch, launch list of the default bench command
python tools/bench_wall_gemm.py 129 4096 > $O/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wall_gemm_sk -s 2 -c 1 -f -o $O/r2_wall_gemm python tools/bench_wall_gemm.py 129 4096 > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log
python bench.py --workload fixed_boundary --steps 1 --no-extras --no-cpu-baseline > $O/plain_fixed.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_picard_resident -s 3 -c 1 -f -o $O/r2_picard_B4096 python bench.py --workload fixed_boundary --steps 1 --no-extras --no-cpu-baseline > $O/ncu_picard.log 2>&1; tail -2 $O/ncu_picard.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; tail -2 $O/ncu_launches.log
ls -la $O | tail -20
