#!/bin/bash
O=gpurun_out
python tools/bench_batch257.py 257 256 10 > $O/plain_b257.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_relax|k_topo$|k_prolong_add|k_jacobi|k_source_scale|k_copy_from_cur' -s 60 -c 8 -f -o $O/r2_stencils python tools/bench_batch257.py 257 256 10 > $O/ncu_st.log 2>&1; tail -2 $O/ncu_st.log
