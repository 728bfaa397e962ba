#!/bin/bash
# Multi-GPU session (N = number of visible GPUs): slab parity over NCCL / peer memory, slab scaling, batch scaling.
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_streaming_parity.py::test_second_device_in_one_process -m gpu -x -q > $O/r2_multi_tests.log 2>&1; tail -5 $O/r2_multi_tests.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  for g in 4097 8193; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --workload slab --grid $g --steps 5 > $O/r2_slab_n${n}_g${g}.json 2> $O/r2_slab_n${n}_g${g}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2_slab_n${n}_g${g}.json").read().strip().splitlines()[-1])
    print("slab n=$n g=$g", "ms/vcycle", round(d["ms_per_vcycle"],3), "glups", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), d["config"]["halo_transport"], "graph", d["config"]["cuda_graph"], "cycles", d["config"]["cycles"], d["config"]["residual_linf"])
except Exception as e:
    print("slab n=$n g=$g FAILED", e)
PY
    tail -2 $O/r2_slab_n${n}_g${g}.err | cut -c1-300
  done
  if [ $n -gt 1 ]; then
    timeout 300 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --workload slab --steps 5 --slab-nccl > $O/r2_slab_nccl_n${n}.json 2> $O/r2_slab_nccl_n${n}.err
    cut -c1-200 $O/r2_slab_nccl_n${n}.json | grep -o '"ms_per_step": [0-9.]*' ; grep -o '"ms_per_vcycle": [0-9.]*' $O/r2_slab_nccl_n${n}.json
  fi
done
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 400 $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --gpus $n --steps 3 --scaling strong --no-extras --no-cpu-baseline > $O/r2_strong_n${n}.json 2> $O/r2_strong_n${n}.err
  grep -o '"value": [0-9.]*' $O/r2_strong_n${n}.json | head -1; tail -1 $O/r2_strong_n${n}.err | cut -c1-200
done
if [ $N -ge 2 ]; then
  timeout 400 $TR --nproc-per-node $N --master-port 29800 bench.py --gpus $N --steps 3 --no-cpu-baseline > $O/r2_weak_n${N}.json 2> $O/r2_weak_n${N}.err
  grep -o '"value": [0-9.]*' $O/r2_weak_n${N}.json | head -3
fi
